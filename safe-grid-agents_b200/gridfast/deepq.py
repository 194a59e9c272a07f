"""Batched deep-Q agent: N lock-step environments, one Q network.

Host object over the sgk_dqn_* entry points (include/sgk.h); replaces
DeepQAgent + ReplayBuffer + dqn_warmup + dqn_learn of the reference
(common/agents/value.py:61-187, common/utils/contain.py:8-22,
common/warmup.py:8-23, common/learn.py:29-58).
"""
import ctypes

import numpy as np
import torch

from ._lib import check
from .batched import _p, _stream


class BatchedDeepQ:
    def __init__(self, env, n_layers=2, n_hidden=100, replay_capacity=10000, batch_size=64,
                 lr=1e-3, discount=0.99, epsilon=0.01, epsilon_anneal=100000, sync_every=10000,
                 reference_bxb_loss=True, seed=0):
        self.L = env.L
        self.env = env
        self.h = ctypes.c_void_p()
        with torch.cuda.device(env.device):
            check(self.L.sgk_dqn_create(env.h, n_layers, n_hidden, replay_capacity, batch_size, seed,
                                        ctypes.byref(self.h)))
        self.n_params = self.L.sgk_dqn_param_count(self.h)
        self.dims = [env.hw] + [n_hidden] * n_layers + [env.n_actions]
        self.batch_size = batch_size
        self.replay_capacity = replay_capacity
        self._added = 0            # transitions appended so far (ReplayBuffer.add calls)
        self.configure(lr, discount, epsilon, epsilon_anneal, sync_every, reference_bxb_loss)

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            self.L.sgk_dqn_destroy(h)

    def configure(self, lr, discount, epsilon, epsilon_anneal, sync_every, reference_bxb_loss=True):
        check(self.L.sgk_dqn_configure(self.h, lr, discount, epsilon, epsilon_anneal, sync_every,
                                       int(reference_bxb_loss)))

    # -- parameters, torch order: (weight [out,in], bias [out]) per Linear -----
    def get_params(self, which=0):
        out = torch.empty(self.n_params, dtype=torch.float32, device=self.env.device)
        check(self.L.sgk_dqn_get_params(self.h, which, _p(out), _stream()))
        return out

    def get_grads(self):
        """Gradients of the last learn step, before clipping (flat, torch order)."""
        out = torch.empty(self.n_params, dtype=torch.float32, device=self.env.device)
        check(self.L.sgk_dqn_get_grads(self.h, _p(out), _stream()))
        return out

    def set_params(self, flat, which=0):
        flat = flat.to(self.env.device, torch.float32).contiguous()
        assert flat.numel() == self.n_params
        check(self.L.sgk_dqn_set_params(self.h, which, _p(flat), _stream()))
        torch.cuda.current_stream().synchronize()

    def load_torch_module(self, module, which=0):
        """Copy the Linear layers of a torch module (e.g. the reference's
        build_Q Sequential, value.py:148-158) into network `which`."""
        flat = torch.cat([p.detach().reshape(-1).float() for m in module.modules()
                          if isinstance(m, torch.nn.Linear) for p in (m.weight, m.bias)])
        self.set_params(flat, which)

    def set_tensor_cores(self, mode=3):
        """3 (or True): tcgen05 with 3xTF32 forwards -- fp32 accuracy, the default
        for the reference's architecture; 1: single-pass TF32; 0 (or False):
        fp32 FFMA kernels."""
        mode = 3 if mode is True else int(mode)
        check(self.L.sgk_dqn_set_tensor_cores(self.h, mode))

    @property
    def tensor_core_mode(self):
        return self.L.sgk_dqn_get_tensor_cores(self.h)

    @property
    def precision(self):
        return {0: "fp32 FFMA", 1: "tcgen05, single-pass TF32 operands, fp32 accumulate",
                3: "tcgen05, 3xTF32 forward (fp32-accurate) + single-pass TF32 backward, fp32 accumulate"}[self.tensor_core_mode]

    def sync_target(self):
        check(self.L.sgk_dqn_sync_target(self.h, _stream()))

    # -- the agent interface, batched --------------------------------------------
    def q_values(self, boards, which=0):
        boards = boards.contiguous()
        n = boards.shape[0]
        out = torch.empty(n, self.env.n_actions, dtype=torch.float32, device=self.env.device)
        check(self.L.sgk_dqn_qvalues(self.h, which, _p(boards), n, _p(out), _stream()))
        return out

    def act(self, boards):
        return self.q_values(boards).argmax(1).to(torch.uint8)

    def replay_add(self, s, a, r, s2, term):
        self._added += s.shape[0]
        check(self.L.sgk_dqn_replay_add(self.h, _p(s), _p(a), _p(r), _p(s2), _p(term), s.shape[0], _stream()))

    def replay_rows(self, first, n):
        """Ring rows [first, first + n): (s u8 [n,HW], a u8 [n], r f32 [n], s2 u8 [n,HW], term u8 [n])."""
        dev = self.env.device
        s = torch.empty(n, self.env.hw, dtype=torch.uint8, device=dev)
        s2 = torch.empty_like(s)
        a = torch.empty(n, dtype=torch.uint8, device=dev)
        term = torch.empty_like(a)
        r = torch.empty(n, dtype=torch.float32, device=dev)
        check(self.L.sgk_dqn_replay_get(self.h, first, n, _p(s), _p(a), _p(r), _p(s2), _p(term), _stream()))
        return s, a, r, s2, term

    @property
    def replay_count(self):
        return self.L.sgk_dqn_replay_count(self.h)

    @property
    def replay_position(self):
        """Ring row the next transition of environment 0 is written to."""
        return self._added % self.replay_capacity

    def last_scalars_device(self):
        out = torch.empty(3, dtype=torch.float32, device=self.env.device)
        check(self.L.sgk_dqn_last_scalars(self.h, _p(out), _stream()))
        return out

    def learn(self, step):
        out = torch.empty(3, dtype=torch.float32, device=self.env.device)
        check(self.L.sgk_dqn_learn(self.h, step, _p(out), _stream()))
        return out

    def learn_batch(self, s, a, r, s2, term):
        """One optimiser step on an explicit batch; returns (loss, grad_norm, clip)."""
        out = torch.empty(3, dtype=torch.float32, device=self.env.device)
        check(self.L.sgk_dqn_learn_batch(self.h, _p(s), _p(a), _p(r), _p(s2), _p(term), s.shape[0], _p(out), _stream()))
        return out

    def warmup(self, n_steps):
        """dqn_warmup: random-policy lock-steps that only fill the ring."""
        check(self.L.sgk_rollout_dqn(self.env.h, self.h, n_steps, self.env.t, 0, _stream()))
        self._added += n_steps * self.env.n
        self.env.t += n_steps

    def rollout(self, n_steps, cheat=False):
        """n_steps lock-steps of act_explore / step / replay.add / learn / sync;
        `cheat` = args.cheat (learn.py:39-47): learn from the hidden reward and
        the action really executed."""
        check(self.L.sgk_rollout_dqn(self.env.h, self.h, n_steps, self.env.t, 1 | (2 if cheat else 0), _stream()))
        self._added += n_steps * self.env.n
        self.env.t += n_steps

    def last_scalars(self):
        out = torch.empty(3, dtype=torch.float32, device=self.env.device)
        check(self.L.sgk_dqn_last_scalars(self.h, _p(out), _stream()))
        return out.tolist()
