from .bleu_evaluator import BLEUEvaluator
from .cls_evaluator import ClsEvaluator
from .evaluator import DatasetEvaluator, DatasetEvaluators, inference_context, inference_on_dataset
from .ppl_evaluator import PPLEvaluator
from .reg_evaluator import RegEvaluator
from .utils import flatten_results_dict, pad_batch, print_csv_format
