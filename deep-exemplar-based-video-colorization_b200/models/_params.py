"""Parameter containers that reproduce the reference's state_dict keys without its nn.Conv2d graph.

The drop-in modules hold plain nn.Parameters (so load_state_dict / .cuda() / .eval() / .parameters()
of test.py:147-166 work) and hand them to libdvc.so; no torch operator ever touches them.
"""
import math

import torch
import torch.nn as nn


class ConvParams(nn.Module):
    """weight [Cout,Cin,k,k] (+ bias [Cout]); initialised like nn.Conv2d (U(-1/sqrt(fan_in), 1/sqrt(fan_in)))."""

    def __init__(self, cin, cout, k=3, bias=True, groups=1):
        super().__init__()
        fan_in = (cin // groups) * k * k
        bound = 1.0 / math.sqrt(fan_in)
        self.weight = nn.Parameter(torch.empty(cout, cin // groups, k, k).uniform_(-bound, bound))
        if bias:
            self.bias = nn.Parameter(torch.empty(cout).uniform_(-bound, bound))


class SlopeParam(nn.Module):
    """nn.PReLU() stand-in: one scalar slope, default 0.25."""

    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.full((1,), 0.25))


def indexed(children):
    """Module whose children are named by integer strings (the gaps in nn.Sequential numbering included)."""
    m = nn.Module()
    for idx, child in children.items():
        m.add_module(str(idx), child)
    return m
