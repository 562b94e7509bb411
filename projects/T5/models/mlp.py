from projects.T5.models.t5_model import *  # noqa: F401,F403  (single-file implementation; see t5_model.py)
