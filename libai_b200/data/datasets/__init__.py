from .bert_dataset import BertDataset
from .cifar import CIFAR10Dataset, CIFAR100Dataset
from .gpt_dataset import GPT2Dataset
from .imagenet import ImageNetDataset
from .mnist import MNISTDataset
from .roberta_dataset import RobertaDataset
from .synthetic import SyntheticBertDataset, SyntheticGPTDataset, SyntheticImageDataset
from .t5_dataset import T5Dataset
