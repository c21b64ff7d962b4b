"""TEST INFRASTRUCTURE ONLY -- stand-in for the un-vendored third-party package `sru`.

The reference imports `from sru import SRU` (/root/reference/src/models/layers/rnn_layers.py:6)
and builds `SRU(input_size=512, hidden_size=32, num_layers=4, bidirectional=True)`
(rnn_layers.py:100-105), calling `self.rnn(x)[0]` on a time-major tensor (rnn_layers.py:150).
The package is pinned as `sru==2.6.0` / git HEAD of taolei87/sru in
/root/reference/setup/requirements.yaml:18,33 and is NOT present in /root/reference nor in
this image, so this file restates its published recurrence.

PARITY UNPINNED for this component: there is no upstream source, test or golden vector to
check the restatement against.  Everything else in the oracle is pinned against the
reference's own Python code executed in the authoring container.

Only the default-argument configuration the reference uses is supported
(dropout=0, use_tanh=False, layer_norm=False, highway_bias=0, has_skip_term=True,
rescale=False, projection_size=0).
"""
import math

import torch
import torch.nn as nn

from oracle.sru_ref import sru_layer_forward


class SRUCell(nn.Module):
    def __init__(self, input_size: int, hidden_size: int, bidirectional: bool = False):
        super().__init__()
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.bidirectional = bidirectional
        self.num_directions = 2 if bidirectional else 1
        self.output_size = hidden_size * self.num_directions
        # k = 4 (extra highway projection) when the skip connection needs a size change
        self.num_matrices = 3 if input_size == self.output_size else 4
        self.weight = nn.Parameter(torch.empty(input_size, self.output_size * self.num_matrices))
        self.weight_c = nn.Parameter(torch.empty(2 * self.output_size))
        self.bias = nn.Parameter(torch.empty(2 * self.output_size))
        self.reset_parameters()

    def reset_parameters(self):
        # upstream init: weight ~ U(+-sqrt(3/in)); weight_c ~ U(+-sqrt(3))*sqrt(.5); bias 0
        val = math.sqrt(3.0 / self.input_size)
        nn.init.uniform_(self.weight, -val, val)
        nn.init.uniform_(self.weight_c, -math.sqrt(3.0), math.sqrt(3.0))
        with torch.no_grad():
            self.weight_c.mul_(math.sqrt(0.5))
        nn.init.zeros_(self.bias)

    def forward(self, x: torch.Tensor):
        return sru_layer_forward(x, self.weight, self.weight_c, self.bias, self.hidden_size, self.bidirectional)


class SRU(nn.Module):
    def __init__(self, input_size, hidden_size, num_layers=2, dropout=0.0, rnn_dropout=0.0, bidirectional=False, **kwargs):
        super().__init__()
        assert dropout == 0.0 and rnn_dropout == 0.0, "oracle shim supports the reference's configuration only"
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.num_layers = num_layers
        self.bidirectional = bidirectional
        out = hidden_size * (2 if bidirectional else 1)
        self.rnn_lst = nn.ModuleList(
            [SRUCell(input_size if i == 0 else out, hidden_size, bidirectional) for i in range(num_layers)]
        )

    def forward(self, x: torch.Tensor, c0=None):
        assert c0 is None
        cs = []
        for cell in self.rnn_lst:
            x, c = cell(x)
            cs.append(c)
        return x, torch.stack(cs)
