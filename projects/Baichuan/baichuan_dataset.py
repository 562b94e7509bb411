"""``BaichuanDataset`` (reference projects/Baichuan/baichuan_dataset.py): pre-tokenised SFT samples."""
from projects.common.sft import SFTDataset


class BaichuanDataset(SFTDataset):
    pass


AlpacaDataset = BaichuanDataset      # the reference's class name for the same pre-tokenised Alpaca samples
