"""``AlpacaDataset`` (reference projects/Llama/dataset.py)."""
from projects.common.sft import SFTDataset


class AlpacaDataset(SFTDataset):
    pass
