from projects.ConvNeXT.modeling.convnext_model import *  # noqa: F401,F403
