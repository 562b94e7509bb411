from .clue_dataset import ClueDataset
from .glue_dataset import GlueDataset

__all__ = ["ClueDataset", "GlueDataset"]
