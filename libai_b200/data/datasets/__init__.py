from .synthetic import SyntheticBertDataset, SyntheticGPTDataset, SyntheticImageDataset
