"""CLUE dataset (reference projects/text_classification/dataset/clue_dataset.py)."""
from .glue_dataset import _TaskDataset
from .utils import Split  # noqa: F401  (the reference defines the enum in this module)
from .utils_clue import clue_output_modes, clue_processors


class ClueDataset(_TaskDataset):
    processors, output_modes = clue_processors, clue_output_modes
