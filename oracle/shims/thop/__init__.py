"""TEST INFRASTRUCTURE ONLY -- stand-in for `thop` (absent): `profile` reports 0 MACs.
Used by /root/reference/src/models/utils/utils.py:5,80-86."""


def profile(model, inputs=(), verbose=False, **kwargs):
    return 0, 0
