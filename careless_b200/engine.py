"""Thin Python owner of one ``clb_handle`` (one GPU, one stream).

This is plumbing between the careless-shaped model classes in ``careless_b200.models`` and
the C-ABI of ``libcareless_b200.so``; all numerics run in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _lib as L


@dataclass
class EngineConfig:
    """Mirror of ``clb_config`` (see include/careless_b200.h for the reference citations)."""
    n_refl: int
    n_meta: int
    mlp_width: int
    mlp_layers: int
    n_refl_total: int = 0
    n_images: int = 0
    image_scales: bool = False
    image_layers: int = 0
    refine_uncertainties: bool = False
    mc_samples: int = 1
    likelihood: str = "normal"
    dof: Optional[float] = None
    laue: bool = False
    prior: str = "wilson"
    n_asu: int = 0
    optimize_dw_r: bool = False
    scale_bijector: str = "exp"
    scale_shift: Optional[float] = None
    epsilon: float = 1e-7
    kl_weight: Optional[float] = None
    learning_rate: float = 1e-3
    beta_1: float = 0.9
    beta_2: float = 0.99
    adam_epsilon: float = 1e-7
    clipnorm: Optional[float] = None
    clipvalue: Optional[float] = None
    global_clipnorm: Optional[float] = None
    seed: int = 1234
    device: int = 0
    stream: int = 0
    rank: int = 0
    world_size: int = 1
    deterministic: bool = False

    def to_c(self) -> L.clb_config:
        c = L.clb_config()
        c.abi_version = L.ABI_VERSION
        c.device = self.device
        c.stream = self.stream or None
        c.n_refl = self.n_refl
        c.n_refl_total = self.n_refl_total or self.n_refl
        c.n_meta, c.mlp_width, c.mlp_layers = self.n_meta, self.mlp_width, self.mlp_layers
        c.n_images = self.n_images if (self.image_scales or self.image_layers > 0) else 0
        c.image_layers = self.image_layers
        c.refine_uncertainties = int(self.refine_uncertainties)
        c.image_scales = int(self.image_scales)
        c.mc_samples = self.mc_samples
        c.likelihood = {"normal": L.LIK_NORMAL, "studentt": L.LIK_STUDENTT}[self.likelihood]
        c.dof = float(self.dof) if self.dof is not None else 0.0
        c.laue = int(self.laue)
        c.prior = {"wilson": L.PRIOR_WILSON, "double_wilson": L.PRIOR_DOUBLE_WILSON}[self.prior]
        c.n_asu = self.n_asu
        c.optimize_dw_r = int(self.optimize_dw_r)
        c.scale_bijector = {"exp": L.BIJ_EXP, "softplus": L.BIJ_SOFTPLUS}[self.scale_bijector]
        c.scale_shift = 0.0 if self.scale_shift is None else float(self.scale_shift)
        c.epsilon = self.epsilon
        c.use_kl_weight = int(self.kl_weight is not None)
        c.kl_weight = 0.0 if self.kl_weight is None else float(self.kl_weight)
        c.learning_rate, c.beta_1, c.beta_2, c.adam_epsilon = self.learning_rate, self.beta_1, self.beta_2, self.adam_epsilon
        c.clipnorm = self.clipnorm or 0.0
        c.clipvalue = self.clipvalue or 0.0
        c.global_clipnorm = self.global_clipnorm or 0.0
        c.seed = self.seed & 0xFFFFFFFFFFFFFFFF
        c.rank, c.world_size = self.rank, self.world_size
        c.deterministic = int(self.deterministic)
        return c


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    def __init__(self, cfg: EngineConfig):
        self.cfg = cfg
        self.lib = L.load()
        self._h = C.c_void_p()
        ccfg = cfg.to_c()
        L.check(self.lib.clb_create(C.byref(ccfg), C.byref(self._h)))
        self.n_rows_total = 0
        self.has_comm = False

    # -- lifetime -------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.clb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        L.check(rc, self._h)

    # -- data -----------------------------------------------------------------------
    def set_observations(self, refl_id, image_id, metadata, intensities, uncertainties,
                         harmonic_id=None, obs_index=None, n_rows_total=0, order=L.ORDER_AUTO):
        refl_id = np.ascontiguousarray(np.asarray(refl_id).reshape(-1), dtype=np.int64)
        n = refl_id.shape[0]
        image_id = None if image_id is None else np.ascontiguousarray(np.asarray(image_id).reshape(-1), dtype=np.int64)
        metadata = np.ascontiguousarray(np.asarray(metadata, dtype=np.float32).reshape(n, -1))
        if metadata.shape[1] != self.cfg.n_meta:
            raise ValueError(f"metadata has {metadata.shape[1]} columns, engine was built for {self.cfg.n_meta}")
        intensities = np.ascontiguousarray(np.asarray(intensities).reshape(-1), dtype=np.float32)
        uncertainties = np.ascontiguousarray(np.asarray(uncertainties).reshape(-1), dtype=np.float32)
        harmonic_id = None if harmonic_id is None else np.ascontiguousarray(np.asarray(harmonic_id).reshape(-1), dtype=np.int64)
        obs_index = None if obs_index is None else np.ascontiguousarray(np.asarray(obs_index).reshape(-1), dtype=np.int64)
        for a in (image_id, intensities, uncertainties, harmonic_id, obs_index):
            if a is not None and a.shape[0] != n:
                raise ValueError("input arrays disagree on the number of rows")
        self.n_rows_total = int(n_rows_total or n)
        self._check(self.lib.clb_set_observations(self._h, n, self.n_rows_total, _ptr(refl_id), _ptr(image_id),
                                                  _ptr(metadata), _ptr(intensities), _ptr(uncertainties),
                                                  _ptr(harmonic_id), _ptr(obs_index), order))

    def download_rows(self):
        """The prepared (sorted / padded SoA) rows as they sit on the device, plus ``ll_const`` and ``prep_ms`` (wall time of
        the row preparation inside the last ``set_observations``): the same dictionary ``clb_prepare_rows`` fills on the host."""
        npad = C.c_int64(); llc = C.c_double(); ms = C.c_double()
        self._check(self.lib.clb_download_rows(self._h, 0, C.byref(npad), None, None, None, None, None, None, None, C.byref(llc), C.byref(ms)))
        m, d = npad.value, self.cfg.n_meta
        out = dict(refl=np.empty(m, np.int32), image=np.zeros(m, np.int32), spot=np.full(m, -1, np.int32), oidx=np.empty(m, np.uint32),
                   meta=np.empty((d, m), np.float32), iobs=np.empty(m, np.float32), sig=np.empty(m, np.float32))
        self._check(self.lib.clb_download_rows(self._h, m, C.byref(npad), _ptr(out["refl"]), _ptr(out["image"]), _ptr(out["spot"]),
                                               _ptr(out["oidx"]), _ptr(out["meta"]), _ptr(out["iobs"]), _ptr(out["sig"]),
                                               C.byref(llc), C.byref(ms)))
        out["ll_const"] = llc.value; out["prep_ms"] = ms.value
        return out

    def download_rows_info(self):
        """``n_padded``, ``ll_const`` and ``prep_ms`` of the prepared rows without copying them."""
        npad = C.c_int64(); llc = C.c_double(); ms = C.c_double()
        self._check(self.lib.clb_download_rows(self._h, 0, C.byref(npad), None, None, None, None, None, None, None, C.byref(llc), C.byref(ms)))
        return {"n_padded": npad.value, "ll_const": llc.value, "prep_ms": ms.value}

    def upload_observations(self):
        self._check(self.lib.clb_upload_observations(self._h))

    def prefetch_observations(self):
        """Asynchronous copy of the prepared rows into the second device buffer; the next step switches to it."""
        self._check(self.lib.clb_prefetch_observations(self._h))

    def set_prior(self, centric, multiplicity, sigma=None, dw_parent=None, asu_id=None, r=None,
                  refl_index=None, init_scale=1.0):
        R = self.cfg.n_refl
        centric = np.ascontiguousarray(np.asarray(centric).reshape(-1).astype(bool), dtype=np.uint8)
        mult = np.ascontiguousarray(np.asarray(multiplicity).reshape(-1), dtype=np.float32)
        if sigma is not None:
            sigma = np.ascontiguousarray(np.broadcast_to(np.asarray(sigma, dtype=np.float32), (R,)))
        dw_parent = None if dw_parent is None else np.ascontiguousarray(dw_parent, dtype=np.int32)
        asu_id = None if asu_id is None else np.ascontiguousarray(asu_id, dtype=np.int32)
        r = None if r is None else np.ascontiguousarray(r, dtype=np.float32)
        refl_index = None if refl_index is None else np.ascontiguousarray(refl_index, dtype=np.int64)
        if centric.shape[0] != R or mult.shape[0] != R:
            raise ValueError("centric / multiplicity must have n_refl entries")
        self._check(self.lib.clb_set_prior(self._h, _ptr(centric), _ptr(mult), _ptr(sigma), _ptr(dw_parent),
                                           _ptr(asu_id), _ptr(r), _ptr(refl_index), float(init_scale)))

    # -- parameters -----------------------------------------------------------------
    def group_size(self, group) -> int:
        return int(self.lib.clb_group_size(self._h, L.GROUPS[group]))

    def get_params(self, group) -> np.ndarray:
        out = np.empty(self.group_size(group), dtype=np.float32)
        self._check(self.lib.clb_get_params(self._h, L.GROUPS[group], _ptr(out), out.size))
        return out

    def set_params(self, group, values):
        v = np.ascontiguousarray(np.asarray(values).reshape(-1), dtype=np.float32)
        self._check(self.lib.clb_set_params(self._h, L.GROUPS[group], _ptr(v), v.size))

    def get_grads(self, group) -> np.ndarray:
        out = np.empty(self.group_size(group), dtype=np.float32)
        self._check(self.lib.clb_get_grads(self._h, L.GROUPS[group], _ptr(out), out.size))
        return out

    def get_adam_state(self, group):
        n = self.group_size(group)
        m, v = np.empty(n, dtype=np.float32), np.empty(n, dtype=np.float32)
        t = C.c_int64()
        self._check(self.lib.clb_get_adam_state(self._h, L.GROUPS[group], _ptr(m), _ptr(v), n, C.byref(t)))
        return m, v, int(t.value)

    def set_trainable(self, group, flag: bool):
        self._check(self.lib.clb_set_trainable(self._h, L.GROUPS[group], int(bool(flag))))

    # -- stepping -------------------------------------------------------------------
    def step(self, n_steps=1, u_f=None, eps_s=None):
        """n full-batch ELBO gradient + Adam steps.  Returns a list of metric dicts (history rows)."""
        S, R, N = self.cfg.mc_samples, self.cfg.n_refl, self.n_rows_total
        if u_f is not None:
            u_f = np.ascontiguousarray(np.asarray(u_f, dtype=np.float32).reshape(n_steps, S, R))
        if eps_s is not None:
            eps_s = np.ascontiguousarray(np.asarray(eps_s, dtype=np.float32).reshape(n_steps, S, N))
        out = (L.clb_metrics * n_steps)()
        done = C.c_int32()
        self._check(self.lib.clb_step(self._h, n_steps, _ptr(u_f), _ptr(eps_s), out, C.byref(done)))
        return [{"loss": m.loss, "NLL": m.nll, "F KLDiv": m.kl, "Grad Norm": m.grad_norm} for m in out[:done.value]]

    def eval(self, u_f=None, eps_s=None):
        """Forward only (keras test_on_batch): metrics of the current parameters on this engine's rows."""
        S, R, N = self.cfg.mc_samples, self.cfg.n_refl, self.n_rows_total
        if u_f is not None:
            u_f = np.ascontiguousarray(np.asarray(u_f, dtype=np.float32).reshape(S, R))
        if eps_s is not None:
            eps_s = np.ascontiguousarray(np.asarray(eps_s, dtype=np.float32).reshape(S, N))
        m = L.clb_metrics()
        self._check(self.lib.clb_eval(self._h, _ptr(u_f), _ptr(eps_s), C.byref(m)))
        return {"loss": m.loss, "NLL": m.nll, "F KLDiv": m.kl}

    def step_begin(self, u_f=None, eps_s=None):
        S, R, N = self.cfg.mc_samples, self.cfg.n_refl, self.n_rows_total
        if u_f is not None:
            u_f = np.ascontiguousarray(np.asarray(u_f, dtype=np.float32).reshape(S, R))
        if eps_s is not None:
            eps_s = np.ascontiguousarray(np.asarray(eps_s, dtype=np.float32).reshape(S, N))
        self._check(self.lib.clb_step_begin(self._h, _ptr(u_f), _ptr(eps_s)))

    def step_norms(self):
        self._check(self.lib.clb_step_norms(self._h))

    def step_end(self, want_metrics=True):
        if not want_metrics:
            self._check(self.lib.clb_step_end(self._h, None))
            return None
        m = L.clb_metrics()
        self._check(self.lib.clb_step_end(self._h, C.byref(m)))
        return {"loss": m.loss, "NLL": m.nll, "F KLDiv": m.kl, "Grad Norm": m.grad_norm}

    def reduce_buffers(self):
        """(ptr_f32, n_f32, ptr_f64, n_f64): device buffers to all-reduce between the step phases."""
        pf, pd = C.c_void_p(), C.c_void_p()
        nf, nd = C.c_int64(), C.c_int64()
        self._check(self.lib.clb_reduce_buffers(self._h, C.byref(pf), C.byref(nf), C.byref(pd), C.byref(nd)))
        return pf.value, int(nf.value), pd.value, int(nd.value)

    def init_comm(self, unique_id: bytes):
        """Library-side NCCL communicator (clb_comm_init): afterwards ``step(n)`` runs on world_size GPUs without Python in the loop."""
        if len(unique_id) != 128:
            raise ValueError("the communicator id is 128 bytes")
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self.lib.clb_comm_init(self._h, buf))
        self.has_comm = True

    # -- debug / measurement ---------------------------------------------------------
    def get_samples(self) -> np.ndarray:
        out = np.empty((self.cfg.mc_samples, self.cfg.n_refl), dtype=np.float32)
        self._check(self.lib.clb_get_samples(self._h, _ptr(out), out.size))
        return out

    def enable_ipred(self, flag=True):
        self._check(self.lib.clb_enable_ipred(self._h, int(flag)))

    def get_ipred(self) -> np.ndarray:
        out = np.empty((self.cfg.mc_samples, self.n_rows_total), dtype=np.float32)
        self._check(self.lib.clb_get_ipred(self._h, _ptr(out), out.size))
        return out

    def get_results(self):
        """Merged F, SigF, I, SigI and redundancy N per surrogate entry (io/manager.py:188-197, :209)."""
        R = self.cfg.n_refl
        out = {k: np.empty(R, dtype=np.float32) for k in ("F", "SigF", "I", "SigI", "N")}
        self._check(self.lib.clb_get_results(self._h, _ptr(out["F"]), _ptr(out["SigF"]), _ptr(out["I"]), _ptr(out["SigI"]),
                                             _ptr(out["N"]), R))
        return out

    def get_scale_moments(self):
        """(mean, stddev) of the scale distribution of every observation, original row order (variational.py:67-69)."""
        n = self.n_rows_total
        mean, std = np.empty(n, dtype=np.float32), np.empty(n, dtype=np.float32)
        self._check(self.lib.clb_get_scale_moments(self._h, _ptr(mean), _ptr(std), n))
        return mean, std

    def synchronize(self):
        self._check(self.lib.clb_synchronize(self._h))

    def reset_timers(self, enable=True):
        self._check(self.lib.clb_reset_timers(self._h, int(enable)))

    def kernel_times(self):
        ms, n, tot = C.c_double(), C.c_int64(), C.c_int64()
        self._check(self.lib.clb_kernel_time_ms(self._h, C.byref(ms), C.byref(n), C.byref(tot)))
        return {"obs_kernel_ms": ms.value, "obs_kernel_launches": int(n.value), "total_launches": int(tot.value)}
