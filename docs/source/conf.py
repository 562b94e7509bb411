import os
import sys

sys.path.insert(0, os.path.abspath("../.."))
project = "libai_b200"
author = "libai_b200 contributors"
extensions = ["myst_parser", "sphinx.ext.autodoc", "sphinx.ext.napoleon", "sphinx.ext.viewcode"]
source_suffix = {".rst": "restructuredtext", ".md": "markdown"}
html_theme = "furo"
autodoc_mock_imports = ["torch", "numpy", "sentencepiece", "transformers", "safetensors", "torchvision"]
