"""Host-side mirror of the reference's two-electron plugin boundary.

Reference (paths relative to /root/reference):
  * `abstract type TwoElectronBuilder` + generic `add_2e_term!`   src/types/ScfProblem.jl:2-9
  * `struct JKBuilderFromTensor` / its `add_2e_term!` method      src/algorithms/JKBuilderFromTensor.jl:7-51

`JKBuilderCUDA` is the drop-in subtype: it owns an immutable ERI (shard) in HBM and its
`add_2e_term` is one call into libscf_b200.so (`scfb_fock_jk`), exactly what the Julia glue
in `julia/JKBuilderCUDA.jl` does with `ccall`.  Same names, argument meaning and error
behaviour as the reference: shape mismatches raise AssertionError (Julia `@assert`),
the abstract fallback raises an error asking to overload.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class TwoElectronBuilder:
    """src/types/ScfProblem.jl:2 — type for objects that build two-electron terms."""

    def add_2e_term(self, out, density, total_density):
        # src/types/ScfProblem.jl:4-9
        raise NotImplementedError("Please overload add_2e_term! for your TwoElectronBuilder type")


class JKBuilderCUDA(TwoElectronBuilder):
    """J+K builder whose ERI lives in B200 HBM.

    n_bas       number of basis functions
    layout      "full" (n^4 fp64, Julia order) or "packed8"
    device      CUDA device ordinal
    devices     list of device ordinals: one builder sharded over several GPUs of this process
    rank,n_ranks  nu-slab shard owned by this builder (one builder per GPU/process)
    """

    def __init__(self, n_bas, layout="full", device=0, rank=0, n_ranks=1, allow_partial=False, devices=None):
        self._h = capi._h()
        self.n_bas = int(n_bas)
        self.layout = layout
        self.device = device
        self.rank, self.n_ranks = rank, n_ranks
        lay = {"full": capi.LAYOUT_FULL, "packed8": capi.LAYOUT_PACKED8}[layout]
        if devices is not None:
            # ONE builder over several GPUs of this process (scfb_create_multi): the single-threaded
            # caller of the reference's run_scf keeps calling add_2e_term once per Fock build
            assert rank == 0 and n_ranks == 1, "devices= and (rank, n_ranks) are mutually exclusive"
            devs = (C.c_int * len(devices))(*[int(d) for d in devices])
            self.device, self.devices = int(devices[0]), tuple(int(d) for d in devices)
            capi.check(capi.lib().scfb_create_multi(C.byref(self._h), self.n_bas, lay, len(devices), devs))
            return
        self.devices = (device,)
        capi.check(capi.lib().scfb_create(C.byref(self._h), self.n_bas, lay, device, rank, n_ranks))
        if allow_partial:
            self.allow_partial()

    # ---- constructors -------------------------------------------------------------
    @classmethod
    def from_tensor(cls, eri, **kw):
        """Equivalent of `JKBuilderFromTensor(eri)` (load_integral_hdf5.jl:15): `eri` is the
        full (n,n,n,n) tensor in Julia index order; it is uploaded once and stays resident."""
        eri = np.asarray(eri)
        assert eri.ndim == 4 and len(set(eri.shape)) == 1, "eri must be (n,n,n,n)"
        self = cls(eri.shape[0], **kw)
        e = capi.as_f64_colmajor(eri)
        capi.check(capi.lib().scfb_eri_upload(self._h, capi.ptr(e)), self._h)
        return self

    def upload_slabs(self, slabs, nu_first):
        """Chunked upload (both layouts): `slabs` is (n,n,n,count) holding the nu-slabs
        nu_first..nu_first+count-1; slabs outside this builder's shard are skipped (packed8: the
        shard is a range of first indices i and row (i,j) is packed on the device from slab i).
        Call `mark_ready()` once the shard is complete (load_integral_hdf5.jl:9-10 TODO: the
        reference wants to avoid holding the whole tensor in host memory)."""
        s = capi.as_f64_colmajor(slabs)
        assert s.ndim == 4 and s.shape[:3] == (self.n_bas,) * 3
        capi.check(capi.lib().scfb_eri_upload_slabs(self._h, capi.ptr(s), int(nu_first), s.shape[3]),
                   self._h)

    def mark_ready(self):
        capi.check(capi.lib().scfb_eri_mark_ready(self._h), self._h)

    @classmethod
    def synthetic(cls, n_bas, family="hash", seed=20260101, aux=None, **kw):
        """ERI generated on the device from a closed-form family."""
        self = cls(n_bas, **kw)
        fam = {"hash": capi.ERI_HASH, "chain": capi.ERI_CHAIN}[family]
        auxp = None
        if aux is not None:
            aux = np.ascontiguousarray(aux, dtype=np.float64)
            auxp = capi.ptr(aux)
        capi.check(capi.lib().scfb_eri_generate(self._h, fam, seed, auxp), self._h)
        return self

    # ---- the boundary -------------------------------------------------------------
    def add_2e_term(self, out, density, total_density):
        """`add_2e_term!(out, builder, density, total_density)`:
        out[:,:,s] += J[total_density] - K[density[:,:,s]]   (JKBuilderFromTensor.jl:22-51).
        `out` is modified in place and returned."""
        n = self.n_bas
        density = np.asarray(density)
        assert density.ndim == 3, "density must be (n_bas, n_bas, n_spin)"
        n_spin = density.shape[2]
        assert n_spin == 1 or n_spin == 2
        assert np.shape(total_density) == (n, n)
        assert density.shape == (n, n, n_spin)
        assert out.shape == (n, n, n_spin)
        d = capi.as_f64_colmajor(density)
        t = capi.as_f64_colmajor(total_density)
        inplace = out.dtype == np.float64 and out.flags.f_contiguous and out.flags.writeable
        o = out if inplace else capi.as_f64_colmajor(out).copy(order="F")
        capi.check(capi.lib().scfb_fock_jk(self._h, n_spin, capi.ptr(d), capi.ptr(t), capi.ptr(o)),
                   self._h)
        if not inplace:
            out[...] = o
        return out

    def jk(self, density, total_density):
        """(J, K) separately: J (n,n), K (n,n,n_spin).  Not in the reference API; used by
        parity tests and the post-SCF energy split (examples/HartreeFock.jl:88-95)."""
        n = self.n_bas
        density = np.asarray(density)
        n_spin = density.shape[2]
        assert n_spin in (1, 2) and density.shape == (n, n, n_spin)
        assert np.shape(total_density) == (n, n)
        d = capi.as_f64_colmajor(density)
        t = capi.as_f64_colmajor(total_density)
        j = np.zeros((n, n), order="F")
        k = np.zeros((n, n, n_spin), order="F")
        capi.check(capi.lib().scfb_jk_split(self._h, n_spin, capi.ptr(d), capi.ptr(t),
                                            capi.ptr(j), capi.ptr(k)), self._h)
        return j, k

    def add_2e_term_device(self, out_ptr, density_ptr, total_density_ptr, n_spin):
        """Device-pointer twin (asynchronous on the builder's stream)."""
        capi.check(capi.lib().scfb_fock_jk_dev(self._h, n_spin, density_ptr, total_density_ptr,
                                               out_ptr), self._h)

    # ---- introspection ------------------------------------------------------------
    @property
    def shard_range(self):
        lo, hi = C.c_int(), C.c_int()
        capi.check(capi.lib().scfb_shard_range(self._h, C.byref(lo), C.byref(hi)), self._h)
        return lo.value, hi.value

    @property
    def eri_bytes(self):
        b = C.c_uint64()
        capi.check(capi.lib().scfb_eri_bytes(self._h, C.byref(b)), self._h)
        return b.value

    def eri_shard(self):
        """Owned nu-slabs copied back to the host, shape (n,n,n,hi-lo) (small n only)."""
        assert self.layout == "full", "use eri_shard_raw() for the packed layout"
        lo, hi = self.shard_range
        n = self.n_bas
        buf = np.empty((n, n, n, hi - lo), order="F")
        capi.check(capi.lib().scfb_eri_download(self._h, capi.ptr(buf)), self._h)
        return buf

    def eri_shard_raw(self):
        """The shard exactly as stored in HBM, flat (packed8: pre-scaled packed rows)."""
        buf = np.empty(self.eri_bytes // 8)
        capi.check(capi.lib().scfb_eri_download(self._h, capi.ptr(buf)), self._h)
        return buf

    def last_build_ms(self):
        a, b = C.c_double(), C.c_double()
        capi.check(capi.lib().scfb_last_build_ms(self._h, C.byref(a), C.byref(b)), self._h)
        return a.value, b.value

    def build_ms_history(self, max_builds=128):
        """(stream_ms[], total_ms[]) of the last builds (<= 128), oldest first; the builds need not
        have been synchronised one by one (bench.py reads this after its back-to-back loop)."""
        a = np.zeros(max_builds)
        b = np.zeros(max_builds)
        cnt = C.c_int()
        capi.check(capi.lib().scfb_build_ms_history(self._h, max_builds, capi.ptr(a), capi.ptr(b), C.byref(cnt)),
                   self._h)
        return a[:cnt.value], b[:cnt.value]

    def syev_jacobi(self, a, v0=None):
        """Eigen-decomposition of one (n, n) or a stack (batch, n, n) of symmetric matrices with the
        small-matrix Jacobi kernels the device-resident SCF step uses (n <= 512); `v0`: orthonormal
        warm-start bases of the same shape.  Returns (w ascending, z with eigenvectors in columns, sweeps)."""
        a = np.asarray(a, dtype=np.float64)
        single = a.ndim == 2
        a3 = a[None] if single else a
        batch, n, _ = a3.shape
        # column-major matrices, consecutive: a symmetric matrix is its own transpose; bases are not
        a_c = np.ascontiguousarray(np.transpose(a3, (0, 2, 1)))
        v_c = None
        if v0 is not None:
            v3 = np.asarray(v0, dtype=np.float64)
            v3 = v3[None] if single else v3
            v_c = np.ascontiguousarray(np.transpose(v3, (0, 2, 1)))
        w = np.zeros((batch, n))
        z = np.zeros((batch, n, n))
        sweeps = (C.c_int * batch)()
        capi.check(capi.lib().scfb_syev_jacobi(self._h, n, batch, capi.ptr(a_c), capi.ptr(v_c) if v_c is not None else None,
                                               capi.ptr(w), capi.ptr(z), sweeps), self._h)
        z = np.transpose(z, (0, 2, 1))  # stored column-major
        sw = list(sweeps)
        return (w[0], z[0], sw[0]) if single else (w, z, sw)

    @property
    def launch_count(self):
        c = C.c_uint64()
        capi.check(capi.lib().scfb_launch_count(self._h, C.byref(c)), self._h)
        return c.value

    def set_stream(self, cuda_stream):
        capi.check(capi.lib().scfb_set_stream(self._h, C.c_void_p(cuda_stream)), self._h)

    def sync(self):
        capi.check(capi.lib().scfb_sync(self._h), self._h)

    def peer_export(self) -> bytes:
        buf = C.create_string_buffer(128)
        capi.check(capi.lib().scfb_peer_export(self._h, buf), self._h)
        return buf.raw

    def peer_import(self, all_handles: bytes):
        assert len(all_handles) == 128 * self.n_ranks
        capi.check(capi.lib().scfb_peer_import(self._h, all_handles), self._h)

    def allow_partial(self, allow=True):
        """A rank of a multi-rank builder without an exchange refuses to build (its result would be
        this rank's partial only); this opts in to exactly that (shard tests / shard timing)."""
        capi.check(capi.lib().scfb_allow_partial(self._h, int(bool(allow))), self._h)

    def comm_init(self, id128: bytes):
        capi.check(capi.lib().scfb_comm_init(self._h, id128), self._h)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            capi.lib().scfb_destroy(self._h)
            self._h = capi._h()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    capi.check(capi.lib().scfb_comm_unique_id(buf))
    return buf.raw
