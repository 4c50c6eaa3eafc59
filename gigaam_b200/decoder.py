"""CTC / RNN-T heads with the reference's interface and state_dict keys (gigaam/decoder.py).  The heads are
parameter holders: their arithmetic (fp32, as in the reference, gigaam/__init__.py:188-189) runs inside the
greedy-decode kernels (`gam_ctc_greedy`, `gam_rnnt_greedy`)."""
from __future__ import annotations

from typing import Dict

import torch

from . import synthetic
from ._params import Bound, build_tree


def _zeros(entries):
    return [(k, torch.zeros(shape, dtype=torch.float32)) for k, shape, kind, _ in entries]


class CTCHead(Bound):
    """gigaam/decoder.py:7-21 -- Conv1d(feat_in, num_classes, k=1) under `decoder_layers.0`."""

    def __init__(self, feat_in: int, num_classes: int):
        super().__init__()
        self.feat_in, self.num_classes = feat_in, num_classes
        build_tree(self, _zeros(synthetic.head_param_list(dict(type="ctc", feat_in=feat_in, num_classes=num_classes))), "head.")


class RNNTHead(Bound):
    """gigaam/decoder.py:140-149 -- `decoder` (Embedding + LSTM) and `joint` (enc / pred / joint_net) holders."""

    def __init__(self, decoder: Dict[str, int], joint: Dict[str, int]):
        super().__init__()
        self.decoder_cfg, self.joint_cfg = dict(decoder), dict(joint)
        build_tree(self, _zeros(synthetic.head_param_list(dict(type="rnnt", decoder=self.decoder_cfg, joint=self.joint_cfg))), "head.")
        self.decoder.blank_id = decoder["num_classes"] - 1
        self.decoder.pred_hidden = decoder["pred_hidden"]
