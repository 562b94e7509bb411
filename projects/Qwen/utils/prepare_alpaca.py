"""Alpaca → tokenised SFT corpus for Qwen; the reference ships this step as ``utils/prepare_alpaca.py``, here it is
the shared ``projects.common.sft.prepare_sft_corpus`` behind ``utils/data_prepare.py`` — this file keeps the name."""
from projects.Qwen.utils.data_prepare import main

if __name__ == "__main__":
    main()
