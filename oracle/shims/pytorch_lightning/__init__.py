"""TEST INFRASTRUCTURE ONLY -- stand-in for `pytorch_lightning` (absent); the model code only reads
`__version__` (/root/reference/src/models/TDAVNet/base_av_model.py:46) and subclasses
LightningModule in the (out-of-scope) video autoencoder."""
import torch.nn as nn

__version__ = "2.1.3"


class LightningModule(nn.Module):
    pass
