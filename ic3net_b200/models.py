"""The reference's independent-controller baselines (``models.py:8-96``: ``MLP``, ``Random``, ``RNN``) on the same
CUDA kernels as the CommNet policy.

They are CommNet steps without communication: ``MLP`` is the non-recurrent tanh step with the comm term switched off
(h = tanh(affine2(x) + x), x = tanh(affine1(obs)), models.py:23-25), ``RNN`` is either an LSTM cell on the encoded
observation (rnn_type LSTM, models.py:76-82) or the vanilla tanh recurrence tanh(affine2(h) + affine1(obs))
(models.py:83-85).  Parameter names and shapes are the reference's (``affine1``, ``affine2``, ``lstm_unit``,
``heads.k``, ``value_head``), so checkpoints interchange; the comm projection the kernels expect is a frozen zero
buffer (``comm_mask_zero`` semantics: S = 0 and the C bias is zero).
"""
import copy

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from .comm import CommNetMLP


class _NoComm(CommNetMLP):
    """CommNet machinery with the communication term frozen at zero."""

    def __init__(self, args, num_inputs):
        a = copy.copy(args)
        a.comm_passes, a.share_weights, a.comm_mask_zero, a.hard_attn = 1, False, True, False
        a.comm_mode, a.comm_init = 'avg', 'zeros'
        if not hasattr(a, 'commnet'):
            a.commnet = False
        super(_NoComm, self).__init__(a, num_inputs)

    def _zero_comm(self, H):
        self.register_buffer('_c_w0', torch.zeros(H, H), persistent=False)
        self.register_buffer('_c_b0', torch.zeros(H), persistent=False)

    def _heads(self, args, H):
        self.heads = nn.ModuleList([nn.Linear(H, o) for o in args.naction_heads])
        self.value_head = nn.Linear(H, 1)

    def _common(self):
        return dict(c_w=[self._c_w0], c_b=[self._c_b0], value_w=self.value_head.weight, value_b=self.value_head.bias,
                    head_w=[h.weight for h in self.heads], head_b=[h.bias for h in self.heads])


class MLP(_NoComm):
    """models.py:8-36: x = tanh(affine1(obs)); h = tanh(affine2(x) + x); heads(h), value_head(h)."""

    def __init__(self, args, num_inputs):
        a = copy.copy(args)
        a.recurrent = False
        super(MLP, self).__init__(a, num_inputs)

    def _variant(self, args):
        return dict(cell=_lib.CELL_TANH, passes=1, x_tanh=1, h_from_x=1)

    def _build_modules(self, args, num_inputs):
        H = args.hid_size
        self.affine1 = nn.Linear(num_inputs, H)
        self.affine2 = nn.Linear(H, H)
        self._heads(args, H)
        self._zero_comm(H)

    def _kernel_weights(self):
        return dict(enc_w=self.affine1.weight, enc_b=self.affine1.bias, f_w=[self.affine2.weight],
                    f_b=[self.affine2.bias], **self._common())


class RNN(_NoComm):
    """models.py:59-96: LSTMCell(affine1(obs), prev_hid) (rnn_type LSTM) or tanh(affine2(prev_hid) + affine1(obs))."""

    def __init__(self, args, num_inputs):
        self.lstm = getattr(args, 'rnn_type', 'MLP') == 'LSTM'
        a = copy.copy(args)
        a.recurrent, a.rnn_type = True, 'LSTM'            # CommNetMLP's own check; the cell is chosen by _variant
        super(RNN, self).__init__(a, num_inputs)
        self.recurrent = True

    def _variant(self, args):
        return dict(cell=_lib.CELL_LSTM if self.lstm else _lib.CELL_TANH, passes=1, x_tanh=0, h_from_x=0)

    def _build_modules(self, args, num_inputs):
        H = args.hid_size
        self.affine1 = nn.Linear(num_inputs, H)
        if self.lstm:
            self.lstm_unit = nn.LSTMCell(H, H)            # models.py:63-65 (affine2 is deleted)
        else:
            self.affine2 = nn.Linear(H, H)
        self._heads(args, H)
        self._zero_comm(H)

    def _kernel_weights(self):
        w = dict(enc_w=self.affine1.weight, enc_b=self.affine1.bias, **self._common())
        if self.lstm:
            w.update(w_ih=self.lstm_unit.weight_ih, w_hh=self.lstm_unit.weight_hh, b_ih=self.lstm_unit.bias_ih,
                     b_hh=self.lstm_unit.bias_hh)
        else:
            w.update(f_w=[self.affine2.weight], f_b=[self.affine2.bias])
        return w

    def forward(self, x, info={}):
        """x = [state, prev_hid] -> (log-probs, value [B, N, 1], next_hid)  (models.py:68-91)."""
        state, hid = x
        B = state.shape[0]
        if not self.lstm and torch.is_tensor(hid):
            hid = hid.reshape(B * self.nagents, self.hid_size)
        action, value, ret = super(RNN, self).forward([state, hid], {})     # independent controllers ignore info
        return action, value.view(B, self.nagents, 1), ret


class Random(nn.Module):
    """models.py:39-57: random log-probabilities and values; nothing is learned."""

    def __init__(self, args, num_inputs):
        super(Random, self).__init__()
        _lib.require_cuda()
        self.naction_heads = args.naction_heads
        self.parameter = nn.Parameter(torch.randn(3, device='cuda'))

    def forward(self, x, info={}):
        sizes = x.size()[:-1]
        v = torch.rand(sizes + (1,), device=x.device)
        out = [F.log_softmax(torch.randn(sizes + (o,), device=x.device), dim=-1) for o in self.naction_heads]
        return out, v
