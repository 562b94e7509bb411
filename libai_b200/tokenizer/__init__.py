from .build import build_tokenizer
from .tokenization_base import PreTrainedTokenizer
from .tokenization_bert import BertTokenizer
from .tokenization_gpt2 import GPT2Tokenizer
from .tokenization_roberta import RobertaTokenizer
from .tokenization_t5 import T5Tokenizer

__all__ = ["build_tokenizer", "PreTrainedTokenizer", "BertTokenizer", "GPT2Tokenizer", "RobertaTokenizer", "T5Tokenizer"]
