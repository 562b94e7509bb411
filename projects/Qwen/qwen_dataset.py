"""``QwenDataset`` (reference projects/Qwen/qwen_dataset.py): pre-tokenised SFT samples."""
from projects.common.sft import SFTDataset


class QwenDataset(SFTDataset):
    pass
