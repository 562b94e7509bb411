from .samplers import CyclicSampler, SingleRoundSampler
