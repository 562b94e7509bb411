from .dataset_utils import (
    compile_helper,
    create_masked_lm_predictions,
    get_samples_mapping,
    get_train_valid_test_split_,
    is_shared_folder,
)
from .indexed_dataset import (
    IndexedCachedDataset,
    IndexedDataset,
    MMapIndexedDataset,
    get_indexed_dataset,
    make_builder,
    make_dataset,
)
