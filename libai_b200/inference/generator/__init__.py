from .generation_beam_search import BeamHypotheses, BeamScorer, BeamSearchScorer
from .generation_logits_processor import *  # noqa: F401,F403
from .generation_stopping_criteria import MaxLengthCriteria, MaxTimeCriteria, StoppingCriteriaList
from .generation_utils import Generator
