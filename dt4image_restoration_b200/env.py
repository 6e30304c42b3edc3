"""Drop-in for the reference environment ``PnPEnv`` (reference ``evaluation/env.py:30-125``).

Call-compatible with the reference's callers (``evaluation/eval.py:75,119,203-212``,
``evaluation/mcts.py:118,126,192,215``): same constructor, ``reset(data, device_type)``,
``step(states, action_dict) -> (states, done)``, ``get_policy_ob``, ``compute_reward``; same state-dict
keys, dtypes, in-place dict mutation and "fresh tensors for x, z, u" ownership.
Differences, all generalisations: any batch ``B`` and any power-of-two ``H, W`` in 32..512 instead of the
literal ``1 x 128 x 128`` (env.py:64,115), and the compute runs as sm_100a CUDA kernels through the C-ABI
(``include/pnp_b200.h``).  There is no CPU path.
"""
from __future__ import annotations

from collections import OrderedDict

import torch

from . import ops


class PnPEnv:
    def __init__(self, max_episode_step, denoiser, device_type) -> None:
        self.max_episode_step = max_episode_step
        self.denoiser = denoiser.to(device_type)
        self.no_ref_model = None
        self._prep_cache = OrderedDict()      # (y0, mask) identity -> ops.ProxPrepared, see _prepared()
        self._load_no_ref()

    def _load_no_ref(self):
        """The reference downloads ARNIQA through torch.hub here (env.py:36-40); that needs the network and
        third-party weights, so the no-reference reward is a plug-in: assign ``env.no_ref_model``."""
        self.no_ref_model = None

    def run_no_ref_reward(self, state):
        if self.no_ref_model is None:
            raise NotImplementedError("no-reference reward model not installed (reference env.py:36-54 uses ARNIQA "
                                      "via torch.hub); assign env.no_ref_model = callable(state) -> float")
        return float(self.no_ref_model(state))

    # ------------------------------------------------------------------------------------------
    def reset(self, data, device_type):
        """env.py:57-71 for any ``[B,1,H,W,2]`` item batch."""
        x = torch.as_tensor(data['x0'])
        x = torch.view_as_complex(x.contiguous())
        data['complex_y0'] = data['y0']
        B, _, H, W = x.shape
        z = x.clone().detach()
        u = torch.zeros_like(x)
        mask = torch.as_tensor(data['mask']).reshape(-1, 1, H, W).contiguous().to(torch.bool)
        y0 = torch.view_as_complex(torch.as_tensor(data['y0']).contiguous())
        gt = torch.as_tensor(data['gt'])
        Aty0 = torch.as_tensor(data['ATy0'])[..., 0]
        x, z, u, mask, y0, gt = (t.to(device_type) for t in (x, z, u, mask, y0, gt))
        return OrderedDict({'x': x, 'y0': y0, 'z': z, 'u': u, 'mask': mask, 'gt': gt, 'ATy0': Aty0, 'T': 0,
                            'complex_y0': data['y0']})

    def _prepared(self, y0, mask):
        """``y0`` and ``mask`` are constants of a trajectory (set in ``reset``), so what the prox step derives from them
        (transposed / column-transformed copies, the mask-structure flag) is prepared once and reused by every later
        ``step`` on the same tensors; the key includes the tensors' version counters, so in-place edits re-prepare."""
        H, W = y0.shape[-2:]
        if not ops.ProxPrepared.supported(H, W):
            return None
        key = (y0.data_ptr(), y0._version, tuple(y0.shape), mask.data_ptr(), mask._version, tuple(mask.shape))
        prep = self._prep_cache.get(key)
        if prep is None:
            prep = ops.ProxPrepared(y0, mask)
            prep._keepalive = (y0, mask)          # the key is only valid while these tensors are alive
            self._prep_cache[key] = prep
            while len(self._prep_cache) > 4:
                self._prep_cache.popitem(last=False)
        else:
            self._prep_cache.move_to_end(key)
        return prep

    def step(self, states: OrderedDict, action_dict: OrderedDict):
        """env.py:74-100: denoise -> centred FFT -> masked k-space solve -> inverse FFT -> dual update."""
        T, mu, sigma_d = action_dict['T'], action_dict['mu'], action_dict['sigma_d']
        y0, z, u, mask = states['y0'], states['z'], states['u'], states['mask']

        if T > 0.5:          # tensor -> bool: same early exit (and host sync) as env.py:79-81
            return states, True
        done = False

        dev = z.device
        mu = torch.as_tensor(mu, dtype=torch.float32, device=dev)
        _mu = mu.view(1, 1, 1, 1)            # scalar mu only, like env.py:88 (RuntimeError otherwise)
        v = ops.residual_real(z, u)          # (z - u).real
        x = self.denoiser(v, torch.as_tensor(sigma_d, dtype=torch.float32, device=dev))
        prep = self._prepared(y0, mask) if (y0.is_cuda and mask.is_cuda) else None
        if prep is not None:
            z, u, _ = prep.prox_dual(x, u, _mu, want_v=False)
        else:
            z, u, _ = ops.prox_dual(x, u, y0, mask, _mu, want_v=False)

        states['x'] = x
        states['z'] = z
        states['u'] = u
        states['T'] = states['T'] + 1 / 30
        return states, done

    @staticmethod
    def get_policy_ob(state: OrderedDict):
        """env.py:103-109; ``[B, H*W]`` (``[1, H*W]`` for the reference's B=1)."""
        x = state['x']
        policy_ob = x.real if x.is_complex() else x
        return policy_ob.reshape(policy_ob.shape[0], -1)

    @staticmethod
    def compute_reward(x, y0):
        """env.py:112-116: PSNR per image as a CPU ``[N,1]`` tensor (computed on the device)."""
        x = x.detach()
        N = x.shape[0]
        gt = y0.detach().reshape(-1, *x.shape[1:]) if y0.numel() != x[0].numel() else y0.detach()
        dev = x.device if x.is_cuda else (gt.device if gt.is_cuda else torch.device('cuda'))
        return ops.psnr(x.to(dev), gt.to(dev)).reshape(N, 1).cpu()


def torch_psnr(output, gt):
    """env.py:120-125 on the device; returns ``[N,1]`` on the inputs' device."""
    return ops.psnr(output, gt).unsqueeze(1)
