"""``AquilaDataset`` (reference projects/Aquila/aquila_dataset.py): pre-tokenised SFT samples."""
from projects.common.sft import SFTDataset


class AquilaDataset(SFTDataset):
    pass
