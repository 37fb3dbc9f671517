"""Host-side owner of one ``libchi`` handle: marshals NumPy (host) or torch-CUDA
(device) buffers through the C ABI.  All arithmetic happens in the CUDA library;
this module only checks shapes, packs parameter blocks and allocates outputs.
"""

import ctypes as C

import numpy as np

from . import _lib, layout


def _f64(a, shape=None, name="array"):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(a.shape)}")
    return a


def _out_f64(out, key, shape):
    """Caller-supplied output array ``out[key]`` (or a fresh one): the library writes ``prod(shape)``
    doubles through its raw pointer, so it must be a C-contiguous float64 array of exactly that shape."""
    a = out.get(key) if out else None
    if a is None:
        return np.empty(shape, dtype=np.float64)
    if not isinstance(a, np.ndarray) or a.dtype != np.float64 or not a.flags.c_contiguous or a.shape != tuple(shape):
        raise ValueError(f"out[{key!r}]: expected a C-contiguous float64 array of shape {tuple(shape)}, got "
                         f"{getattr(a, 'dtype', type(a))} {getattr(a, 'shape', '')}")
    return a


def _sightline(sightline, B):
    if sightline is None:
        return None
    sl = np.ascontiguousarray(sightline, dtype=np.int32)
    if sl.shape != (B,):
        raise ValueError(f"sightline: expected [{B}], got {sl.shape}")
    return sl


class PinnedArray:
    """A NumPy view of page-locked host memory from ``chi_host_alloc`` (fast copies)."""

    def __init__(self, shape, dtype=np.float64):
        lib = _lib.load()
        self._lib = lib
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = C.c_void_p()
        rc = lib.chi_host_alloc(C.byref(ptr), max(nbytes, 1))
        if rc != 0:
            raise _lib.ChiError(rc, "chi_host_alloc failed")
        self._ptr = ptr
        buf = (C.c_char * max(nbytes, 1)).from_address(ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._ptr is not None:
            self.array = None
            self._lib.chi_host_free(self._ptr)
            self._ptr = None

    def __del__(self):  # pragma: no cover
        try:
            self.free()
        except Exception:
            pass


class Engine:
    """One handle on one B200.  ``kind`` is a key of ``layout.KINDS``."""

    def __init__(
        self,
        kind,
        n_clouds,
        *,
        device=0,
        lorentzian=False,
        free_weight=True,
        has_emission=None,
        has_absorption=None,
        bg_temp=3.77,
        depth_nth_fwhm_power=1.0 / 3.0,
        prior_fwhm2=500.0,
        prior_velocity=(0.0, 10.0),
        prior_log10_nHI=(0.0, 1.5),
        prior_log10_n_alpha=(-6.0, 2.0),
        prior_fwhm_L=0.0,
        prior_amplitude=1.0,
    ):
        if kind not in layout.KINDS:
            raise ValueError(f"unknown model kind {kind!r}")
        self._lib = _lib.load()
        self._h = None
        self.kind = kind
        self.n_clouds = int(n_clouds)
        self.lorentzian = bool(lorentzian)
        self.free_weight = bool(free_weight)
        if kind == "physics":
            self.has_emission = True if has_emission is None else bool(has_emission)
            self.has_absorption = True if has_absorption is None else bool(has_absorption)
        else:
            self.has_emission = layout.has_emission(kind)
            self.has_absorption = layout.has_absorption(kind)
        cfg = _lib.ChiConfig()
        cfg.model = layout.KINDS[kind]
        cfg.n_clouds = self.n_clouds
        cfg.has_emission = int(self.has_emission)
        cfg.has_absorption = int(self.has_absorption)
        cfg.lorentzian = int(self.lorentzian)
        cfg.free_weight = int(self.free_weight)
        cfg.dtype = 0
        cfg.bg_temp = float(bg_temp)
        cfg.depth_nth_fwhm_power = float(depth_nth_fwhm_power)
        cfg.prior_fwhm2 = float(prior_fwhm2)
        cfg.prior_velocity = (C.c_double * 2)(*map(float, prior_velocity))
        cfg.prior_log10_nHI = (C.c_double * 2)(*map(float, prior_log10_nHI))
        cfg.prior_log10_n_alpha = (C.c_double * 2)(*map(float, prior_log10_n_alpha))
        cfg.prior_fwhm_L = float(prior_fwhm_L or 0.0)
        cfg.prior_amplitude = float(prior_amplitude)
        h = C.c_void_p()
        rc = self._lib.chi_create(C.byref(h), int(device), C.byref(cfg))
        if rc != 0:
            raise _lib.ChiError(rc, self._lib.chi_last_error(None).decode())
        self._h = h
        self.device = int(device)
        self.slots = layout.slot_names(kind, self.lorentzian, self.free_weight)
        self.S_e = self.S_a = 0
        self.nb_e = self.nb_a = 0
        self.n_sightlines = 0
        self._keep = []

    # ---- lifetime -----------------------------------------------------------
    def close(self):
        if self._h is not None:
            self._lib.chi_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise _lib.ChiError(rc, self._lib.chi_last_error(self._h).decode())

    # ---- spectra ------------------------------------------------------------
    def _dataset(self, d, n_sl):
        ds = _lib.ChiDataset()
        if d is None:
            return ds, 0, 0
        spectral = _f64(d["spectral"])
        S = spectral.shape[0]
        brightness = _f64(d["brightness"]).reshape(-1, S)
        if brightness.shape[0] != n_sl:
            raise ValueError("brightness: expected [n_sightlines, S]")
        # scalar, [S] or [n_sightlines, S] (SpecData accepts a scalar or per-channel noise)
        noise = _f64(np.broadcast_to(np.asarray(d["noise"], dtype=np.float64), (n_sl, S)), (n_sl, S), "noise")
        basis = d.get("basis")
        nb = 0
        if basis is not None:
            basis = _f64(basis).reshape(-1, S)
            nb = basis.shape[0]
        offset = d.get("offset")
        if offset is not None:
            offset = _f64(np.broadcast_to(np.asarray(offset, dtype=np.float64), (S,)))
        ds.n_channels = S
        ds.n_baseline = nb
        ds.spectral = _lib.as_dp(spectral)
        ds.brightness = _lib.as_dp(brightness)
        ds.noise = _lib.as_dp(noise)
        ds.basis = _lib.as_dp(basis) if nb else None
        ds.offset = _lib.as_dp(offset) if offset is not None else None
        self._keep = getattr(self, "_keep", []) + [spectral, brightness, noise, basis, offset]
        return ds, S, nb

    def set_spectra(self, emission=None, absorption=None, n_sightlines=1):
        """Upload observed spectra.  Each dataset is a dict with ``spectral [S]``,
        ``brightness [S]`` or ``[n_sightlines, S]``, ``noise`` (scalar, ``[S]`` or
        ``[n_sightlines, S]``), optional ``basis [D+1, S]`` and ``offset`` (scalar or
        ``[S]``) describing the affine baseline."""
        self._keep = []
        de, self.S_e, self.nb_e = self._dataset(emission if self.has_emission else None, n_sightlines)
        da, self.S_a, self.nb_a = self._dataset(absorption if self.has_absorption else None, n_sightlines)
        self._check(self._lib.chi_set_spectra(self._h, int(n_sightlines), C.byref(de), C.byref(da)))
        self.n_sightlines = int(n_sightlines)
        self._keep = []

    # ---- packing ------------------------------------------------------------
    def pack(self, rvs, B=None):
        """dict of RV arrays (``[B, N]``, ``[N]``, ``[B, 1]`` or ``[B]`` for a
        per-draw scalar) -> ``theta [B, NSLOT, N]``."""
        N = self.n_clouds
        if B is None:
            B = max((np.asarray(v).shape[0] if np.asarray(v).ndim == 2 else 1) for v in rvs.values())
        theta = np.zeros((B, layout.NSLOT, N), dtype=np.float64)
        scal = layout.scalar_rvs(self.kind)
        if self.kind == "physics":
            theta[:, layout.PHYSICS_SLOTS["filling_factor"], :] = 1.0
            theta[:, layout.PHYSICS_SLOTS["absorption_weight"], :] = 1.0
        for name, slot in self.slots.items():
            if name not in rvs:
                if self.kind == "physics":
                    continue
                raise KeyError(f"missing free RV {name!r} for model kind {self.kind!r}")
            v = np.asarray(rvs[name], dtype=np.float64)
            if v.ndim == 1:
                # 1-D: one value per draw for the per-draw scalar RVs, else one per cloud
                v = v[:, None] if name in scal else v[None, :]
            theta[:, slot, :] = np.broadcast_to(v, (B, N))
        return theta

    def unpack_grad(self, grad):
        """``grad [B, NSLOT, N]`` -> dict keyed by RV name; per-draw scalar RVs are
        summed over clouds (``[B, 1]``)."""
        out = {}
        scal = layout.scalar_rvs(self.kind)
        for name, slot in self.slots.items():
            g = grad[:, slot, :]
            out[name] = g.sum(axis=1, keepdims=True) if name in scal else g
        return out

    # ---- model-level, host buffers -------------------------------------------
    def _coeffs(self, coeff, nb, B, name):
        if coeff is None or nb == 0:
            return None
        return _f64(np.broadcast_to(np.asarray(coeff, dtype=np.float64), (B, nb)), (B, nb), name)

    def logp_grad(self, theta, coeff_e=None, coeff_a=None, sightline=None, want_grad=True, out=None):
        """Likelihood logp ``[B]`` and gradients for ``theta [B, NSLOT, N]``.
        Returns ``(logp, grad, gcoeff_e, gcoeff_a)`` (``None`` where not applicable).
        ``out`` may supply preallocated ``logp``/``grad``/``gcoeff_e``/``gcoeff_a`` arrays; with
        page-locked inputs and outputs (:class:`PinnedArray`) the library overlaps the copies of
        consecutive sub-batches with the kernels."""
        N = self.n_clouds
        theta = _f64(theta)
        if theta.ndim != 3 or theta.shape[1:] != (layout.NSLOT, N):
            raise ValueError(f"theta: expected [B, {layout.NSLOT}, {N}], got {theta.shape}")
        B = theta.shape[0]
        ce = self._coeffs(coeff_e, self.nb_e, B, "coeff_e")
        ca = self._coeffs(coeff_a, self.nb_a, B, "coeff_a")
        sl = _sightline(sightline, B)
        logp = _out_f64(out, "logp", (B,))
        grad = gce = gca = None
        if want_grad:
            grad = _out_f64(out, "grad", (B, layout.NSLOT, N))
            gce = _out_f64(out, "gcoeff_e", (B, self.nb_e)) if self.nb_e else None
            gca = _out_f64(out, "gcoeff_a", (B, self.nb_a)) if self.nb_a else None
        self._check(
            self._lib.chi_logp_grad(
                self._h, B, _lib.as_dp(theta), _lib.as_dp(ce), _lib.as_dp(ca), _lib.as_ip(sl),
                _lib.as_dp(logp), _lib.as_dp(grad), _lib.as_dp(gce), _lib.as_dp(gca),
            )
        )
        return logp, grad, gce, gca

    # ---- packed rows: only the free RVs the model kind has ---------------------
    @property
    def row_slots(self):
        """Slots of the model kind's free RVs in ascending order: the rows of the packed
        blocks ``[B, len(row_slots), N]`` taken by :meth:`logp_grad_rows`."""
        return np.array(sorted(self.slots.values()), dtype=np.int32)

    def pack_rows(self, rvs, B=None):
        """Like :meth:`pack` but ``[B, n_rows, N]`` with only the rows in :attr:`row_slots`."""
        return np.ascontiguousarray(self.pack(rvs, B)[:, self.row_slots, :])

    def unpack_rows(self, rows):
        """``rows [B, n_rows, N]`` (values or gradients) -> dict keyed by RV name; per-draw scalar
        RVs are summed over clouds."""
        scal = layout.scalar_rvs(self.kind)
        index = {int(s): r for r, s in enumerate(self.row_slots)}
        out = {}
        for name, slot in self.slots.items():
            g = rows[:, index[slot], :]
            out[name] = g.sum(axis=1, keepdims=True) if name in scal else g
        return out

    def logp_grad_rows(self, rows, coeff_e=None, coeff_a=None, sightline=None, want_grad=True, out=None):
        """:meth:`logp_grad` on packed blocks (C ABI ``chi_logp_grad_rows``): ``rows`` and the
        returned gradient are ``[B, len(row_slots), N]`` -- 10-40 % fewer bytes over PCIe."""
        N = self.n_clouds
        slots = self.row_slots
        rows = _f64(rows)
        if rows.ndim != 3 or rows.shape[1:] != (slots.size, N):
            raise ValueError(f"rows: expected [B, {slots.size}, {N}], got {rows.shape}")
        B = rows.shape[0]
        ce = self._coeffs(coeff_e, self.nb_e, B, "coeff_e")
        ca = self._coeffs(coeff_a, self.nb_a, B, "coeff_a")
        sl = _sightline(sightline, B)
        logp = _out_f64(out, "logp", (B,))
        grad = gce = gca = None
        if want_grad:
            grad = _out_f64(out, "grad", (B, slots.size, N))
            gce = _out_f64(out, "gcoeff_e", (B, self.nb_e)) if self.nb_e else None
            gca = _out_f64(out, "gcoeff_a", (B, self.nb_a)) if self.nb_a else None
        self._check(
            self._lib.chi_logp_grad_rows(
                self._h, B, int(slots.size), _lib.as_ip(slots), _lib.as_dp(rows), _lib.as_dp(ce), _lib.as_dp(ca),
                _lib.as_ip(sl), _lib.as_dp(logp), _lib.as_dp(grad), _lib.as_dp(gce), _lib.as_dp(gca),
            )
        )
        return logp, grad, gce, gca

    def logp_grad_rows_async(self, rows, coeff_e, coeff_a, out, sightline=None):
        """Asynchronous :meth:`logp_grad_rows` (C ABI ``chi_logp_grad_rows_async``): enqueues the
        call and returns a ticket for :meth:`wait`.  ``rows``, the coefficients and every array of
        ``out`` (``logp``, ``grad`` and, with baselines, ``gcoeff_e`` / ``gcoeff_a``) must be
        :class:`PinnedArray` views that stay untouched until the wait returns.  Keeping two calls
        in flight hides one call's pipeline fill and drain behind the other's kernels."""
        N = self.n_clouds
        slots = self.row_slots
        if rows.dtype != np.float64 or not rows.flags.c_contiguous or rows.shape[1:] != (slots.size, N):
            raise ValueError(f"rows: expected a C-contiguous float64 [B, {slots.size}, {N}] array")
        B = rows.shape[0]
        for name, a, shape in (("coeff_e", coeff_e, (B, self.nb_e)), ("coeff_a", coeff_a, (B, self.nb_a)),
                               ("logp", out.get("logp"), (B,)), ("grad", out.get("grad"), rows.shape),
                               ("gcoeff_e", out.get("gcoeff_e"), (B, self.nb_e)),
                               ("gcoeff_a", out.get("gcoeff_a"), (B, self.nb_a))):
            if shape[-1] == 0 and name != "logp":
                continue
            if a is None or a.dtype != np.float64 or not a.flags.c_contiguous or a.shape != shape:
                raise ValueError(f"{name}: expected a C-contiguous float64 array of shape {shape}")
        sl = _sightline(sightline, B)
        ticket = C.c_int64(-1)
        self._check(
            self._lib.chi_logp_grad_rows_async(
                self._h, B, int(slots.size), _lib.as_ip(slots), _lib.as_dp(rows),
                _lib.as_dp(coeff_e if self.nb_e else None), _lib.as_dp(coeff_a if self.nb_a else None), _lib.as_ip(sl),
                _lib.as_dp(out["logp"]), _lib.as_dp(out["grad"]),
                _lib.as_dp(out.get("gcoeff_e") if self.nb_e else None),
                _lib.as_dp(out.get("gcoeff_a") if self.nb_a else None), C.byref(ticket),
            )
        )
        self._keep = (rows, coeff_e, coeff_a, sl, out)  # the arrays of the newest call stay referenced
        return int(ticket.value)

    def logp_sums_rows(self, rows, coeff_e=None, coeff_a=None, sightline=None, out=None, want_logp=True):
        """Totals instead of per-draw gradients (C ABI ``chi_logp_sums_rows``): returns ``(logp [B] | None,
        sums [2 + n_rows N])`` with ``sums = {sum logp, sum of the gradient rows, non-finite draws}`` over
        the finite draws; :meth:`unpack_sums` splits it.  The per-draw gradient never leaves the device."""
        N = self.n_clouds
        slots = self.row_slots
        rows = _f64(rows)
        if rows.ndim != 3 or rows.shape[1:] != (slots.size, N):
            raise ValueError(f"rows: expected [B, {slots.size}, {N}], got {rows.shape}")
        B = rows.shape[0]
        ce = self._coeffs(coeff_e, self.nb_e, B, "coeff_e")
        ca = self._coeffs(coeff_a, self.nb_a, B, "coeff_a")
        sl = _sightline(sightline, B)
        logp = _out_f64(out, "logp", (B,)) if want_logp else None
        sums = _out_f64(out, "sums", (2 + slots.size * N,))
        self._check(self._lib.chi_logp_sums_rows(self._h, B, int(slots.size), _lib.as_ip(slots), _lib.as_dp(rows),
                                                 _lib.as_dp(ce), _lib.as_dp(ca), _lib.as_ip(sl), _lib.as_dp(logp),
                                                 _lib.as_dp(sums)))
        return logp, sums

    def logp_sums_rows_async(self, rows, coeff_e, coeff_a, out, sightline=None):
        """Asynchronous :meth:`logp_sums_rows` (``chi_logp_sums_rows_async``): pinned ``rows`` / coefficients,
        ``out["sums"]`` (any host array) and optionally pinned ``out["logp"]``; the sums are complete after
        :meth:`wait` on the returned ticket."""
        N = self.n_clouds
        slots = self.row_slots
        if rows.dtype != np.float64 or not rows.flags.c_contiguous or rows.shape[1:] != (slots.size, N):
            raise ValueError(f"rows: expected a C-contiguous float64 [B, {slots.size}, {N}] array")
        B = rows.shape[0]
        for name, a, shape in (("coeff_e", coeff_e, (B, self.nb_e)), ("coeff_a", coeff_a, (B, self.nb_a))):
            if shape[-1] and (a is None or a.dtype != np.float64 or not a.flags.c_contiguous or a.shape != shape):
                raise ValueError(f"{name}: expected a C-contiguous float64 array of shape {shape}")
        sums = _out_f64(out, "sums", (2 + slots.size * N,))
        logp = out.get("logp")
        if logp is not None:
            logp = _out_f64(out, "logp", (B,))
        sl = _sightline(sightline, B)
        ticket = C.c_int64(-1)
        self._check(self._lib.chi_logp_sums_rows_async(
            self._h, B, int(slots.size), _lib.as_ip(slots), _lib.as_dp(rows), _lib.as_dp(coeff_e if self.nb_e else None),
            _lib.as_dp(coeff_a if self.nb_a else None), _lib.as_ip(sl), _lib.as_dp(logp), _lib.as_dp(sums), C.byref(ticket)))
        self._keep = (rows, coeff_e, coeff_a, sl, out, sums)
        return int(ticket.value)

    def unpack_sums(self, sums):
        """``sums [2 + n_rows N]`` -> ``(sum logp, dict of summed gradients keyed by RV name, non-finite draws)``."""
        n_rows = self.row_slots.size
        g = np.asarray(sums[1:-1]).reshape(1, n_rows, self.n_clouds)
        return float(sums[0]), {k: v[0] for k, v in self.unpack_rows(g).items()}, int(round(float(sums[-1])))

    def stage_count(self):
        """Number of logp / VJP kernel pipelines (cloud map -> moments -> epilogue) issued so far on
        this handle; :meth:`stage_ms` holds the timings of the latest 64."""
        return int(self._lib.chi_stage_count(self._h))

    def wait(self, ticket=-1):
        """Block until the asynchronous call ``ticket`` (default: everything) has completed."""
        self._check(self._lib.chi_wait(self._h, int(ticket)))

    def spectra(self, theta, coeff_e=None, coeff_a=None, dtype=np.float64):
        """Predicted spectra ``(mu_e [B, S_e] | None, mu_a [B, S_a] | None)``.  ``dtype=np.float32``
        selects the single-precision kernels (``chi_spectra_f32``: float32 outputs within 1e-5 of the
        largest value of each fp64 spectrum; parameters and their maps stay fp64)."""
        dtype = np.dtype(dtype)
        if dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise ValueError("spectra: dtype must be float64 or float32")
        f32 = dtype == np.dtype(np.float32)
        N = self.n_clouds
        theta = _f64(theta)
        B = theta.shape[0]
        if theta.shape[1:] != (layout.NSLOT, N):
            raise ValueError("theta: bad shape")
        ce = self._coeffs(coeff_e, self.nb_e, B, "coeff_e")
        ca = self._coeffs(coeff_a, self.nb_a, B, "coeff_a")
        mu_e = np.empty((B, self.S_e), dtype=dtype) if self.has_emission else None
        mu_a = np.empty((B, self.S_a), dtype=dtype) if self.has_absorption else None
        if f32:
            self._check(
                self._lib.chi_spectra_f32(self._h, B, _lib.as_dp(theta), _lib.as_dp(ce), _lib.as_dp(ca),
                                          _lib.as_fp(mu_e), _lib.as_fp(mu_a))
            )
        else:
            self._check(
                self._lib.chi_spectra(self._h, B, _lib.as_dp(theta), _lib.as_dp(ce), _lib.as_dp(ca),
                                      _lib.as_dp(mu_e), _lib.as_dp(mu_a))
            )
        return mu_e, mu_a

    def predictive(self, theta, coeff_e=None, coeff_a=None, seed=0, draw_offset=0, sightline=None, dtype=np.float64):
        """Predictive draws of the observed nodes, ``y = mu + noise * z`` (``chi_predictive``): what
        ``sample_posterior_predictive`` / ``sample_prior_predictive`` return for the likelihood nodes.
        ``z`` is keyed by ``(seed, draw_offset + b, dataset, channel)``, so shards and sub-batches that
        pass their offset reproduce one big call.  Returns ``(y_e | None, y_a | None)``."""
        N = self.n_clouds
        theta = _f64(theta)
        B = theta.shape[0]
        if theta.shape[1:] != (layout.NSLOT, N):
            raise ValueError("theta: bad shape")
        dtype = np.dtype(dtype)
        if dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise ValueError("predictive: dtype must be float64 or float32")
        ce = self._coeffs(coeff_e, self.nb_e, B, "coeff_e")
        ca = self._coeffs(coeff_a, self.nb_a, B, "coeff_a")
        sl = None if sightline is None else np.ascontiguousarray(sightline, dtype=np.int32)
        if sl is not None and sl.shape != (B,):
            raise ValueError("sightline: expected [B]")
        y_e = np.empty((B, self.S_e), dtype=dtype) if self.has_emission else None
        y_a = np.empty((B, self.S_a), dtype=dtype) if self.has_absorption else None
        if dtype == np.dtype(np.float32):
            fn, cast = self._lib.chi_predictive_f32, _lib.as_fp
        else:
            fn, cast = self._lib.chi_predictive, _lib.as_dp
        self._check(fn(self._h, B, _lib.as_dp(theta), _lib.as_dp(ce), _lib.as_dp(ca), _lib.as_ip(sl),
                       int(seed) & 0xFFFFFFFFFFFFFFFF, int(draw_offset), cast(y_e), cast(y_a)))
        return y_e, y_a

    def predictive_dev(self, B, theta, y_e=None, y_a=None, coeff_e=None, coeff_a=None, seed=0, draw_offset=0,
                       sightline=None, f32=False):
        """:meth:`predictive` on device pointers (``chi_predictive_dev`` / ``chi_predictive_f32_dev``)."""
        p = self._ptr
        fn = self._lib.chi_predictive_f32_dev if f32 else self._lib.chi_predictive_dev
        self._check(fn(self._h, int(B), p(theta), p(coeff_e), p(coeff_a), p(sightline),
                       int(seed) & 0xFFFFFFFFFFFFFFFF, int(draw_offset), p(y_e), p(y_a)))

    def spectra_vjp(self, theta, gbar_e=None, gbar_a=None, coeff_e=None, coeff_a=None):
        """VJP of :meth:`spectra`: returns ``(grad, gcoeff_e, gcoeff_a)``."""
        N = self.n_clouds
        theta = _f64(theta)
        B = theta.shape[0]
        ce = self._coeffs(coeff_e, self.nb_e, B, "coeff_e")
        ca = self._coeffs(coeff_a, self.nb_a, B, "coeff_a")
        ge = _f64(gbar_e, (B, self.S_e), "gbar_e") if (gbar_e is not None and self.has_emission) else None
        ga = _f64(gbar_a, (B, self.S_a), "gbar_a") if (gbar_a is not None and self.has_absorption) else None
        grad = np.empty((B, layout.NSLOT, N), dtype=np.float64)
        gce = np.empty((B, self.nb_e), dtype=np.float64) if self.nb_e else None
        gca = np.empty((B, self.nb_a), dtype=np.float64) if self.nb_a else None
        self._check(
            self._lib.chi_spectra_vjp(self._h, B, _lib.as_dp(theta), _lib.as_dp(ce), _lib.as_dp(ca),
                                      _lib.as_dp(ge), _lib.as_dp(ga), _lib.as_dp(grad), _lib.as_dp(gce),
                                      _lib.as_dp(gca))
        )
        return grad, gce, gca

    # ---- total model logp in PyMC's unconstrained space ------------------------------
    def set_priors(self, beta=(2.0, 2.0), nth_fwhm_1pc=(1.75, 0.25), sigma_log10_NHI=0.5,
                   baseline_sigma_e=None, baseline_sigma_a=None):
        """Hyper-parameters of the prior densities that are not already part of the parameter
        maps: Beta(alpha, beta) of the thermal factor, TruncatedNormal(mu, sigma, lower=0) of
        ``nth_fwhm_1pc``, the width of the coupled Normal of ``log10_wt_ff_*`` and the Normal
        widths of the baseline coefficients (default 1)."""
        p = _lib.ChiPriors()
        p.beta = (C.c_double * 2)(*map(float, beta))
        p.nth_fwhm_1pc = (C.c_double * 2)(*map(float, nth_fwhm_1pc))
        p.sigma_log10_NHI = float(sigma_log10_NHI or 0.0)
        for dst, src in ((p.baseline_sigma_e, baseline_sigma_e), (p.baseline_sigma_a, baseline_sigma_a)):
            vals = [1.0] * _lib.CHI_MAX_BASELINE
            if src is not None:
                vals[: len(src)] = [float(v) for v in src]
            for i, v in enumerate(vals):
                dst[i] = v
        self._check(self._lib.chi_set_priors(self._h, C.byref(p)))

    def model_logp_grad(self, u, coeff_e=None, coeff_a=None, sightline=None, want_grad=True):
        """Total model logp (likelihood + priors + transform Jacobians) and its gradient with
        respect to the UNCONSTRAINED block ``u [B, NSLOT, N]`` -- what PyMC's
        ``logp_dlogp_function`` evaluates inside NUTS.  Returns ``(logp, grad, gcoeff_e, gcoeff_a)``."""
        N = self.n_clouds
        u = _f64(u)
        if u.ndim != 3 or u.shape[1:] != (layout.NSLOT, N):
            raise ValueError(f"u: expected [B, {layout.NSLOT}, {N}], got {u.shape}")
        B = u.shape[0]
        ce = self._coeffs(coeff_e, self.nb_e, B, "coeff_e")
        ca = self._coeffs(coeff_a, self.nb_a, B, "coeff_a")
        sl = None if sightline is None else np.ascontiguousarray(sightline, dtype=np.int32)
        logp = np.empty(B, dtype=np.float64)
        grad = np.empty((B, layout.NSLOT, N), dtype=np.float64) if want_grad else None
        gce = np.empty((B, self.nb_e), dtype=np.float64) if (want_grad and self.nb_e) else None
        gca = np.empty((B, self.nb_a), dtype=np.float64) if (want_grad and self.nb_a) else None
        self._check(
            self._lib.chi_model_logp_grad(self._h, B, _lib.as_dp(u), _lib.as_dp(ce), _lib.as_dp(ca), _lib.as_ip(sl),
                                          _lib.as_dp(logp), _lib.as_dp(grad), _lib.as_dp(gce), _lib.as_dp(gca))
        )
        return logp, grad, gce, gca

    def model_logp_grad_dev(self, B, u, logp, grad=None, coeff_e=None, coeff_a=None, sightline=None,
                            gcoeff_e=None, gcoeff_a=None):
        p = self._ptr
        self._check(
            self._lib.chi_model_logp_grad_dev(self._h, int(B), p(u), p(coeff_e), p(coeff_a), p(sightline),
                                              p(logp), p(grad), p(gcoeff_e), p(gcoeff_a))
        )

    def nuts_sample(self, init_u, init_ce=None, init_ca=None, *, n_tune=1000, n_draws=1000, seed=1234,
                    target_accept=0.8, max_treedepth=10, init_step=0.0, sightline=None, pool_mass=False, lock_step=False):
        """Chain-batched NUTS on the device.  ``init_u [chains, NSLOT, N]`` are unconstrained
        starting points with finite model logp.  Returns a dict with ``u [n_draws, chains, NSLOT,
        N]`` (unconstrained draws), ``coeff_e``/``coeff_a`` and the sampler statistics
        ``accept``, ``depth``, ``diverging``, ``logp``, ``step_size`` (``[n_draws, chains]``)."""
        N = self.n_clouds
        init_u = _f64(init_u)
        chains = init_u.shape[0]
        if init_u.shape[1:] != (layout.NSLOT, N):
            raise ValueError("init_u: bad shape")
        ce = self._coeffs(np.zeros((chains, self.nb_e)) if init_ce is None else init_ce, self.nb_e, chains, "init_ce")
        ca = self._coeffs(np.zeros((chains, self.nb_a)) if init_ca is None else init_ca, self.nb_a, chains, "init_ca")
        sl = None if sightline is None else np.ascontiguousarray(sightline, dtype=np.int32)
        cfg = _lib.ChiNutsConfig(chains=chains, n_tune=int(n_tune), n_draws=int(n_draws),
                                 max_treedepth=int(max_treedepth), pool_mass=int(bool(pool_mass)), lock_step=int(bool(lock_step)), seed=int(seed),
                                 target_accept=float(target_accept), init_step=float(init_step))
        out = {
            "u": np.empty((n_draws, chains, layout.NSLOT, N)),
            "coeff_e": np.empty((n_draws, chains, self.nb_e)) if self.nb_e else None,
            "coeff_a": np.empty((n_draws, chains, self.nb_a)) if self.nb_a else None,
            "accept": np.empty((n_draws, chains)),
            "depth": np.empty((n_draws, chains), dtype=np.int32),
            "diverging": np.empty((n_draws, chains), dtype=np.int32),
            "logp": np.empty((n_draws, chains)),
            "step_size": np.empty((n_draws, chains)),
        }
        self._check(
            self._lib.chi_nuts_sample(
                self._h, C.byref(cfg), _lib.as_dp(init_u), _lib.as_dp(ce), _lib.as_dp(ca), _lib.as_ip(sl),
                _lib.as_dp(out["u"]), _lib.as_dp(out["coeff_e"]), _lib.as_dp(out["coeff_a"]),
                _lib.as_dp(out["accept"]), _lib.as_ip(out["depth"]), _lib.as_ip(out["diverging"]),
                _lib.as_dp(out["logp"]), _lib.as_dp(out["step_size"]),
            )
        )
        return out

    def transform(self, block, direction=1):
        """PyMC's default transforms on ``[B, NSLOT, N]``: ``direction=+1`` unconstrained ->
        constrained, ``-1`` the inverse."""
        block = _f64(block)
        out = np.empty_like(block)
        self._check(self._lib.chi_transform(self._h, block.shape[0], int(direction), _lib.as_dp(block), _lib.as_dp(out)))
        return out

    def deterministics(self, theta):
        """dict of the reference's per-cloud Deterministics, each ``[B, N]``."""
        theta = _f64(theta)
        B = theta.shape[0]
        det = np.empty((B, _lib.CHI_NDET, self.n_clouds), dtype=np.float64)
        self._check(self._lib.chi_deterministics(self._h, B, _lib.as_dp(theta), _lib.as_dp(det)))
        return {name: det[:, i, :] for i, name in enumerate(layout.DET_NAMES)}

    # ---- model-level, device buffers (torch CUDA tensors or raw pointers) ----------
    @staticmethod
    def _ptr(t):
        if t is None:
            return None
        if isinstance(t, int):
            return C.c_void_p(t)
        return C.c_void_p(t.data_ptr())

    def logp_grad_dev(self, B, theta, logp, grad=None, coeff_e=None, coeff_a=None, sightline=None,
                      gcoeff_e=None, gcoeff_a=None):
        """Enqueue on the handle's stream; arguments are device tensors/pointers."""
        p = self._ptr
        self._check(
            self._lib.chi_logp_grad_dev(self._h, int(B), p(theta), p(coeff_e), p(coeff_a), p(sightline),
                                        p(logp), p(grad), p(gcoeff_e), p(gcoeff_a))
        )

    def logp_grad_rows_dev(self, B, rows, logp, grad=None, coeff_e=None, coeff_a=None, sightline=None,
                           gcoeff_e=None, gcoeff_a=None):
        """:meth:`logp_grad_dev` on packed device blocks ``[B, len(row_slots), N]``."""
        p = self._ptr
        slots = self.row_slots
        self._check(
            self._lib.chi_logp_grad_rows_dev(self._h, int(B), int(slots.size), _lib.as_ip(slots), p(rows), p(coeff_e),
                                             p(coeff_a), p(sightline), p(logp), p(grad), p(gcoeff_e), p(gcoeff_a))
        )

    def spectra_dev(self, B, theta, mu_e=None, mu_a=None, coeff_e=None, coeff_a=None):
        p = self._ptr
        self._check(self._lib.chi_spectra_dev(self._h, int(B), p(theta), p(coeff_e), p(coeff_a), p(mu_e), p(mu_a)))

    def spectra_f32_dev(self, B, theta, mu_e=None, mu_a=None, coeff_e=None, coeff_a=None):
        """:meth:`spectra_dev` with float32 outputs (``chi_spectra_f32_dev``)."""
        p = self._ptr
        self._check(self._lib.chi_spectra_f32_dev(self._h, int(B), p(theta), p(coeff_e), p(coeff_a), p(mu_e), p(mu_a)))

    def spectra_vjp_dev(self, B, theta, grad, gbar_e=None, gbar_a=None, coeff_e=None, coeff_a=None,
                        gcoeff_e=None, gcoeff_a=None):
        p = self._ptr
        self._check(
            self._lib.chi_spectra_vjp_dev(self._h, int(B), p(theta), p(coeff_e), p(coeff_a), p(gbar_e),
                                          p(gbar_a), p(grad), p(gcoeff_e), p(gcoeff_a))
        )

    def synchronize(self):
        self._check(self._lib.chi_synchronize(self._h))

    @property
    def stream(self):
        return self._lib.chi_stream(self._h)

    # ---- physics-level ---------------------------------------------------------
    def spin_temp(self, kinetic_temp, density, n_alpha):
        tk, de, na = np.broadcast_arrays(*(np.asarray(x, dtype=np.float64) for x in (kinetic_temp, density, n_alpha)))
        shape = tk.shape
        tk, de, na = (_f64(x).ravel() for x in (tk, de, na))
        out = np.empty_like(tk)
        self._check(self._lib.chi_spin_temp(self._h, tk.size, _lib.as_dp(tk), _lib.as_dp(de), _lib.as_dp(na), _lib.as_dp(out)))
        return out.reshape(shape)

    def spin_temp_vjp(self, kinetic_temp, density, n_alpha, gbar):
        tk, de, na, gb = np.broadcast_arrays(
            *(np.asarray(x, dtype=np.float64) for x in (kinetic_temp, density, n_alpha, gbar))
        )
        shape = tk.shape
        tk, de, na, gb = (_f64(x).ravel() for x in (tk, de, na, gb))
        g = [np.empty_like(tk) for _ in range(3)]
        self._check(
            self._lib.chi_spin_temp_vjp(self._h, tk.size, _lib.as_dp(tk), _lib.as_dp(de), _lib.as_dp(na),
                                        _lib.as_dp(gb), *(_lib.as_dp(x) for x in g))
        )
        return tuple(x.reshape(shape) for x in g)

    ELEMENTWISE = {"gaussian": 0, "lorentzian": 1, "log10_column_density": 2, "log10_nonthermal_pressure": 3}

    def elementwise(self, fn, a, b, c=None):
        """``physics.gaussian/lorentzian(x, center, fwhm)``, ``calc_log10_column_density
        (tau_total, spin_temp)`` and ``calc_log10_nonthermal_pressure(log10_density,
        fwhm2_nonthermal)`` with NumPy broadcasting, evaluated on the device."""
        args = [np.asarray(x, dtype=np.float64) for x in ((a, b) if c is None else (a, b, c))]
        args = np.broadcast_arrays(*args)
        shape = args[0].shape
        flat = [_f64(x).ravel() for x in args]
        out = np.empty_like(flat[0])
        self._check(
            self._lib.chi_elementwise(self._h, self.ELEMENTWISE[fn], flat[0].size, _lib.as_dp(flat[0]),
                                      _lib.as_dp(flat[1]), _lib.as_dp(flat[2]) if c is not None else None,
                                      _lib.as_dp(out))
        )
        return out.reshape(shape)

    def pseudo_voigt(self, velo_axis, velocity, fwhm2, fwhm_L):
        v = _f64(velo_axis)
        c = _f64(np.atleast_1d(velocity))
        N = c.shape[0]
        w = _f64(np.broadcast_to(np.asarray(fwhm2, dtype=np.float64), (N,)))
        fl = _f64(np.broadcast_to(np.asarray(fwhm_L, dtype=np.float64), (N,)))
        out = np.empty((v.shape[0], N), dtype=np.float64)
        self._check(
            self._lib.chi_pseudo_voigt(self._h, v.shape[0], N, _lib.as_dp(v), _lib.as_dp(c), _lib.as_dp(w),
                                       _lib.as_dp(fl), _lib.as_dp(out))
        )
        return out

    def radiative_transfer(self, tau, tspin, filling_factor, bg_temp):
        tau = _f64(tau)
        S, N = tau.shape
        ts = _f64(np.broadcast_to(np.asarray(tspin, dtype=np.float64), (N,)))
        ff = _f64(np.broadcast_to(np.asarray(filling_factor, dtype=np.float64), (N,)))
        out = np.empty(S, dtype=np.float64)
        self._check(
            self._lib.chi_radiative_transfer(self._h, S, N, _lib.as_dp(tau), _lib.as_dp(ts), _lib.as_dp(ff),
                                             float(bg_temp), _lib.as_dp(out))
        )
        return out

    def pseudo_voigt_vjp(self, velo_axis, velocity, fwhm2, fwhm_L, gbar):
        """Gradients of ``sum(gbar * pseudo_voigt)`` w.r.t. velocity, fwhm2, fwhm_L (``[N]``)."""
        v = _f64(velo_axis)
        c = _f64(np.atleast_1d(velocity))
        N = c.shape[0]
        w = _f64(np.broadcast_to(np.asarray(fwhm2, dtype=np.float64), (N,)))
        fl = _f64(np.broadcast_to(np.asarray(fwhm_L, dtype=np.float64), (N,)))
        gb = _f64(gbar, (v.shape[0], N), "gbar")
        g = [np.empty(N, dtype=np.float64) for _ in range(3)]
        self._check(
            self._lib.chi_pseudo_voigt_vjp(self._h, v.shape[0], N, _lib.as_dp(v), _lib.as_dp(c), _lib.as_dp(w),
                                           _lib.as_dp(fl), _lib.as_dp(gb), *(_lib.as_dp(x) for x in g))
        )
        return tuple(g)

    def radiative_transfer_vjp(self, tau, tspin, filling_factor, bg_temp, gbar):
        """Gradients of ``sum(gbar * radiative_transfer)`` w.r.t. tau ``[S, N]``, tspin and
        filling_factor (``[N]``)."""
        tau = _f64(tau)
        S, N = tau.shape
        ts = _f64(np.broadcast_to(np.asarray(tspin, dtype=np.float64), (N,)))
        ff = _f64(np.broadcast_to(np.asarray(filling_factor, dtype=np.float64), (N,)))
        gb = _f64(gbar, (S,), "gbar")
        g_tau = np.empty((S, N), dtype=np.float64)
        g_ts = np.empty(N, dtype=np.float64)
        g_ff = np.empty(N, dtype=np.float64)
        self._check(
            self._lib.chi_radiative_transfer_vjp(self._h, S, N, _lib.as_dp(tau), _lib.as_dp(ts), _lib.as_dp(ff),
                                                 float(bg_temp), _lib.as_dp(gb), _lib.as_dp(g_tau),
                                                 _lib.as_dp(g_ts), _lib.as_dp(g_ff))
        )
        return g_tau, g_ts, g_ff

    # ---- measurement -------------------------------------------------------------
    def last_kernel_ms(self):
        ms = C.c_float()
        self._check(self._lib.chi_last_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def stage_ms(self, n=1):
        """Kernel milliseconds ``[n, 4]`` = (map, emission, absorption, epilogue) of the
        ``n`` most recent logp/VJP batches (oldest first; the library keeps 64)."""
        ms = (C.c_float * (4 * n))()
        self._check(self._lib.chi_stage_ms(self._h, int(n), ms))
        return np.array(ms, dtype=np.float64).reshape(n, 4)

    def launch_count(self):
        return int(self._lib.chi_launch_count(self._h))

    # ---- multi-GPU: the library's own communicator (chi_comm_*, csrc/comm.cuh) -----------------
    @staticmethod
    def comm_unique_id():
        """128 bytes identifying a new communicator (``chi_comm_unique_id``): rank 0 creates it, the caller
        ships it to every rank (e.g. ``torch.distributed.broadcast_object_list``), every rank calls
        :meth:`comm_init`."""
        lib = _lib.load()
        buf = C.create_string_buffer(_lib.CHI_COMM_ID_BYTES)
        rc = lib.chi_comm_unique_id(buf)
        if rc != 0:
            raise _lib.ChiError(rc, lib.chi_last_error(None).decode())
        return bytes(buf.raw)

    def comm_init(self, unique_id, rank, n_ranks):
        if len(unique_id) != _lib.CHI_COMM_ID_BYTES:
            raise ValueError("unique_id: expected the 128 bytes of Engine.comm_unique_id()")
        self._check(self._lib.chi_comm_init(self._h, C.c_char_p(unique_id), int(rank), int(n_ranks)))

    def comm_destroy(self):
        self._check(self._lib.chi_comm_destroy(self._h))

    @property
    def comm_size(self):
        return int(self._lib.chi_comm_size(self._h))

    @property
    def comm_rank(self):
        return int(self._lib.chi_comm_rank(self._h))

    def shard_sums_dev(self, B, width, logp, grad, sums):
        """``sums[0]`` = sum of ``logp`` over the finite draws, ``sums[1:1+width]`` = sum of their gradient
        rows, ``sums[1+width]`` = number of non-finite draws (``chi_shard_sums_dev``; device tensors)."""
        p = self._ptr
        self._check(self._lib.chi_shard_sums_dev(self._h, int(B), int(width), p(logp), p(grad), p(sums)))

    def allgather_dev(self, send, recv, count):
        p = self._ptr
        self._check(self._lib.chi_allgather_dev(self._h, p(send), p(recv), int(count)))

    def allreduce_sum_dev(self, buf, count):
        self._check(self._lib.chi_allreduce_sum_dev(self._h, self._ptr(buf), int(count)))

    @property
    def comm_stream(self):
        """The library's communication stream (``cudaStream_t`` as int) for ``torch.cuda.ExternalStream``."""
        return int(self._lib.chi_comm_stream(self._h) or 0)

    def comm_join(self):
        self._check(self._lib.chi_comm_join(self._h))

    def comm_synchronize(self):
        self._check(self._lib.chi_comm_synchronize(self._h))

    def dev_join(self):
        """``dev_overlap`` mode: order the outputs of the device-resident calls issued so far into the handle's
        stream (``chi_dev_join``)."""
        self._check(self._lib.chi_dev_join(self._h))

    def set_option(self, name, value):
        """Tuning / test knobs (``chi_set_option``): ``small_path`` 0 by batch size, 1 never, 2 always
        (the one-launch evaluation of small batches); ``small_nl`` clouds per thread of that kernel;
        ``small_chunks`` its blocks per draw; ``nuts_tick`` 0 by model size, 1 / 2 graph of kernels per tick
        (global-memory / staged tick), 3 / 4 / 5 / 6 the per-chain kernel (256 / 512 threads / a cluster of blocks per chain / speculative); ``dev_overlap``, ``ea_one_pass``, ``small_cluster``: ``include/chi.h``."""
        opts = {"small_path": _lib.OPT_SMALL_PATH, "small_nl": _lib.OPT_SMALL_NL, "small_chunks": _lib.OPT_SMALL_CHUNKS,
                "small_stamps": _lib.OPT_SMALL_STAMPS, "nuts_tick": _lib.OPT_NUTS_TICK,
                "dev_overlap": _lib.OPT_DEV_OVERLAP, "ea_one_pass": _lib.OPT_EA_ONE_PASS, "small_cluster": _lib.OPT_SMALL_CLUSTER}
        self._check(self._lib.chi_set_option(self._h, opts[name], int(value)))

    def microbench_fp32(self):
        """``(FP32 FMA TFLOP/s, 1e12 MUFU.EX2 per second)`` of this GPU (``chi_microbench_fp32``)."""
        a, b = C.c_double(), C.c_double()
        self._check(self._lib.chi_microbench_fp32(self._h, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def microbench_fp64(self):
        v = C.c_double()
        self._check(self._lib.chi_microbench_fp64(self._h, C.byref(v)))
        return float(v.value)
