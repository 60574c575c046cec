"""CommNet / IC3Net policy on one B200, behind the reference's ``comm.CommNetMLP``
surface (comm.py:8-253): ``CommNetMLP(args, num_inputs)``, ``forward(x, info={})``,
``init_hidden(batch_size)`` and the same ``state_dict`` keys
(``heads.k.*``, ``encoder.*``, ``hidd_encoder.*``, ``f_module.{weight,bias}_{ih,hh}``,
``C_modules.0.*``, ``value_head.*``), so checkpoints interchange.

The forward pass runs the hand-written kernels of csrc/policy.cu through the C ABI
(encoder, gated hidden-state mean, C, LSTMCell, value/action heads) in float32 for a
whole batch ``[B, N, .]`` of environments.  Only the recurrent (LSTM) branch with
``comm_passes == 1`` is accelerated -- the branch every BASELINE config uses; other
variants raise.  The rollout forward is inference-only (the reference detaches what
it samples from, action_utils.py:35); gradients are taken by the trainer.
"""
import ctypes as C

import numpy as np
import torch
from torch import nn

from . import _lib


def _as_u8(v, B, N, device):
    """info['comm_action'] / info['alive_mask'] (numpy or tensor, [N] or [B,N]) -> uint8 [B,N]."""
    if torch.is_tensor(v) and v.is_cuda:
        t = v if v.dtype == torch.uint8 else (v != 0).to(torch.uint8)
    else:       # host mask: normalise on the host, then ONE copy (asynchronous when the source is pinned)
        t = v if torch.is_tensor(v) else torch.from_numpy(np.ascontiguousarray(np.asarray(v)))
        if t.dtype != torch.uint8:
            t = (t != 0).to(torch.uint8)
        t = t.to(device, non_blocking=True)
    if t.dim() == 1:
        t = t.unsqueeze(0).expand(B, N)
    return t.reshape(B, N).contiguous()      # the kernels test `!= 0`, any non-zero byte counts as 1


class CommNetMLP(nn.Module):
    def __init__(self, args, num_inputs):
        super(CommNetMLP, self).__init__()
        _lib.require_cuda()
        self.args = args
        self.nagents = args.nagents
        self.hid_size = args.hid_size
        self.comm_passes = args.comm_passes
        self.recurrent = args.recurrent
        self.continuous = args.continuous
        if self.continuous:
            raise NotImplementedError("continuous actions are outside the accelerated path")
        if not args.recurrent or getattr(args, 'rnn_type', 'LSTM') != 'LSTM':
            raise NotImplementedError("only the recurrent LSTM CommNet/IC3Net branch is accelerated")
        if args.comm_passes != 1 or getattr(args, 'share_weights', False):
            raise NotImplementedError("comm_passes != 1 / share_weights are outside the accelerated path")
        self.num_inputs = num_inputs
        H = args.hid_size
        # parameters with the reference's names and shapes (comm.py:31-96); init like nn.Linear / nn.LSTMCell
        self.heads = nn.ModuleList([nn.Linear(H, o) for o in args.naction_heads])
        self.encoder = nn.Linear(num_inputs, H)
        self.hidd_encoder = nn.Linear(H, H)          # allocated but unused by the reference forward (comm.py:57,125)
        self.f_module = nn.LSTMCell(H, H)
        self.C_modules = nn.ModuleList([nn.Linear(H, H)])
        if args.comm_init == 'zeros':
            self.C_modules[0].weight.data.zero_()
        self.value_head = nn.Linear(H, 1)
        self.to(torch.device('cuda', torch.cuda.current_device()), torch.float32)

        heads = list(args.naction_heads)
        self._atot = sum(heads)
        hd = (C.c_int32 * _lib.MAX_HEADS)(*(heads + [0] * (_lib.MAX_HEADS - len(heads))))
        self._cfg_proto = dict(N=self.nagents, H=H, O=num_inputs, nheads=len(heads), head_dim=hd,
                               hard_attn=int(bool(args.hard_attn)),
                               comm_avg=int(getattr(args, 'comm_mode', 'avg') == 'avg'),
                               comm_mask_zero=int(bool(args.comm_mask_zero)),
                               env_id0=int(getattr(args, 'env_id0', 0)),
                               seed=int(getattr(args, 'seed', 0)) & 0xFFFFFFFFFFFFFFFF,
                               obs_off=0, obs_vocab=0, obs_ncount=0)
        self._packed = None
        self._packed_key = None
        # 'tc' = tcgen05 tensor-core path (csrc/policy_tc.cu, hid_size 128), 'simt' = fp32 CUDA-core kernel
        self.policy_impl = getattr(args, 'policy_impl', None) or ('tc' if H == 128 else 'simt')
        if self.policy_impl not in ('tc', 'simt'):
            raise ValueError("policy_impl must be 'tc' or 'simt'")
        if self.policy_impl == 'tc' and H != 128:
            raise NotImplementedError("the tensor-core policy path is specialised for hid_size 128")
        self._ws = {}

    def set_obs_layout(self, off, vocab, ncount):
        """Observation layout hint (include/ic3net_b200.h, ic3_policy_cfg.obs_vocab): lets the encoder sum the
        one-hot class terms separately from the counts, so the class part can come from a per-position table and
        the dense / index / fused encoders stay bit-identical.  (0, 0, 0) = plain single sum."""
        self._cfg_proto.update(obs_off=int(off), obs_vocab=int(vocab), obs_ncount=int(ncount))

    # ---- kernel-side weights ---------------------------------------------------
    def policy_cfg(self, B):
        return _lib.PolicyCfg(B=B, **self._cfg_proto)

    def workspace(self, B):
        """(scratch, err) tensors of the tensor-core path for a batch of B envs (None, None for 'simt')."""
        if self.policy_impl != 'tc':
            return None, None
        if B not in self._ws:
            cfg = self.policy_cfg(B)
            nbytes = int(_lib.load().ic3_policy_workspace_bytes(C.byref(cfg)))
            dev = self.encoder.weight.device
            self._ws[B] = (torch.empty(nbytes, dtype=torch.uint8, device=dev),
                           torch.zeros(1, dtype=torch.int32, device=dev))
        return self._ws[B]

    def _param_list(self):
        ps = [self.encoder.weight, self.encoder.bias, self.C_modules[0].weight, self.C_modules[0].bias,
              self.f_module.weight_ih, self.f_module.weight_hh, self.f_module.bias_ih, self.f_module.bias_hh,
              self.value_head.weight, self.value_head.bias]
        for hd in self.heads:
            ps += [hd.weight, hd.bias]
        return ps

    def packed(self):
        """K-major kernel layout of the parameters; re-packed (one kernel) when any parameter changed."""
        ps = self._param_list()
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if self._packed is not None and key == self._packed_key:
            return self._packed
        H, O = self.hid_size, self.num_inputs
        dev = self.encoder.weight.device
        if self._packed is None:
            nout = 1 + self._atot
            self._bufs = dict(enc_wT=torch.empty(O, H, device=dev), enc_b=torch.empty(H, device=dev),
                              c_wT=torch.empty(H, H, device=dev), c_b=torch.empty(H, device=dev),
                              lstm_wT=torch.empty(2 * H, 4 * H, device=dev), lstm_b=torch.empty(4 * H, device=dev),
                              head_w=torch.empty(nout, H, device=dev), head_b=torch.empty(nout, device=dev))
            if self.policy_impl == 'tc':
                self._bufs['lstm_img'] = torch.empty(_lib.LSTM_IMG_BYTES, dtype=torch.uint8, device=dev)
                self._bufs['bias_cat'] = torch.empty(4 * H, device=dev)
                self._bufs['flags'] = torch.zeros(1, dtype=torch.int32, device=dev)
            self._packed = _lib.PolicyPacked(**{k: v.data_ptr() for k, v in self._bufs.items()})
        for p in ps:
            assert p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()
        hw = (C.c_void_p * _lib.MAX_HEADS)(*([h.weight.data_ptr() for h in self.heads] +
                                             [None] * (_lib.MAX_HEADS - len(self.heads))))
        hb = (C.c_void_p * _lib.MAX_HEADS)(*([h.bias.data_ptr() for h in self.heads] +
                                             [None] * (_lib.MAX_HEADS - len(self.heads))))
        params = _lib.PolicyParams(encoder_w=ps[0].data_ptr(), encoder_b=ps[1].data_ptr(), c_w=ps[2].data_ptr(),
                                   c_b=ps[3].data_ptr(), w_ih=ps[4].data_ptr(), w_hh=ps[5].data_ptr(),
                                   b_ih=ps[6].data_ptr(), b_hh=ps[7].data_ptr(), value_w=ps[8].data_ptr(),
                                   value_b=ps[9].data_ptr(), head_w=hw, head_b=hb)
        cfg = self.policy_cfg(1)
        _lib.check(_lib.load().ic3_policy_pack(C.byref(cfg), C.byref(params), C.byref(self._packed), _lib.stream()))
        self._packed_key = key
        return self._packed

    # ---- reference surface -------------------------------------------------------
    def forward(self, x, info={}):
        """x = [state [B,N,O], (h, c) each [B*N, H]]; info may hold 'comm_action' and
        'alive_mask' ([N] or [B,N]).  Returns (list of log-probs [B,N,na_k],
        value [B*N,1], (h', c')) like comm.py:134-244."""
        state, (h, c) = x
        B, N, H = state.shape[0], self.nagents, self.hid_size
        dev = state.device
        state = state.to(torch.float32).contiguous()
        h = h.detach().to(dev, torch.float32).contiguous()
        c = c.detach().to(dev, torch.float32).contiguous()
        cfg = self.policy_cfg(B)
        w = self.packed()
        lib = _lib.load()
        xenc = torch.empty(B * N, H, device=dev)
        _lib.check(lib.ic3_encoder_dense(C.byref(cfg), C.byref(w), state.data_ptr(), xenc.data_ptr(), _lib.stream()))
        comm = alive = None
        if self.args.hard_attn:
            comm = _as_u8(info['comm_action'], B, N, dev)          # comm.py:171-175
        if 'alive_mask' in info:
            alive = _as_u8(info['alive_mask'], B, N, dev)          # comm.py:102-104
        h2, c2 = torch.empty_like(h), torch.empty_like(c)
        value = torch.empty(B * N, 1, device=dev)
        logp = torch.empty(B, N, self._atot, device=dev)
        ws, err = self.workspace(B)
        io = _lib.PolicyIO(x=xenc.data_ptr(), h=h.data_ptr(), c=c.data_ptr(), comm_action=_lib.ptr(comm),
                           alive=_lib.ptr(alive), fresh=None, tick=None, draws=None, h_out=h2.data_ptr(),
                           c_out=c2.data_ptr(), value=value.data_ptr(), logp=logp.data_ptr(), action=None,
                           workspace=_lib.ptr(ws), err=_lib.ptr(err))
        _lib.check(lib.ic3_policy_step(C.byref(cfg), C.byref(w), C.byref(io), _lib.stream()))
        action = list(torch.split(logp, list(self.args.naction_heads), dim=-1))
        return action, value, (h2, c2)

    def init_hidden(self, batch_size):
        dev = self.encoder.weight.device
        return tuple((torch.zeros(batch_size * self.nagents, self.hid_size, device=dev),
                      torch.zeros(batch_size * self.nagents, self.hid_size, device=dev)))
