"""TEST INFRASTRUCTURE ONLY -- minimal stand-in for `timm` (absent from this image) so that the
reference's src/models can be imported for golden-vector generation.  Only `DropPath`
(/root/reference/src/models/layers/conv_layers.py:5, attention.py:6) is needed."""
