"""ofhe_b200 -- Python (ctypes) binding of libofhe_b200.so, the sm_100a RNS/NTT backend for OpenFHE's DCRTPoly
hot path.  The binding only marshals pointers: device memory, streams and torch.distributed come from PyTorch
(plumbing), every arithmetic operation runs in the hand-written CUDA kernels behind include/ofhe_b200.h.

There is NO CPU fallback: importing works anywhere (so the CPU-only test tier can check the ABI), but creating a
``Context`` without a CUDA device, or with the shared library missing, raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OFHE_B200_LIB") or os.path.join(os.path.dirname(_HERE), "libofhe_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "include", "ofhe_b200.h")

_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)
_vp = C.c_void_p

# name -> (restype, argtypes): mirrors include/ofhe_b200.h one to one (tests/test_abi.py checks the two agree)
SIGNATURES = {
    "ofhe_last_error": (C.c_char_p, []),
    "ofhe_version": (C.c_char_p, []),
    "ofhe_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "ofhe_ctx_destroy": (None, [_vp]),
    "ofhe_ctx_sm_count": (C.c_int, [_vp]),
    "ofhe_malloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "ofhe_free": (C.c_int, [_vp, _vp]),
    "ofhe_memcpy_h2d": (C.c_int, [_vp, _vp, _vp, C.c_size_t, _vp]),
    "ofhe_memcpy_d2h": (C.c_int, [_vp, _vp, _vp, C.c_size_t, _vp]),
    "ofhe_memcpy_d2d": (C.c_int, [_vp, _vp, _vp, C.c_size_t, _vp]),
    "ofhe_memset0": (C.c_int, [_vp, _vp, C.c_size_t, _vp]),
    "ofhe_stream_sync": (C.c_int, [_vp, _vp]),
    "ofhe_stream_create": (C.c_int, [_vp, C.POINTER(_vp)]),
    "ofhe_stream_destroy": (C.c_int, [_vp, _vp]),
    "ofhe_ring_create": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, _vp, C.POINTER(_vp)]),
    "ofhe_ring_destroy": (None, [_vp]),
    "ofhe_ring_twiddles_host": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp, _u64p, _u64p]),
    "ofhe_ntt_fwd": (C.c_int, [_vp, _vp, _vp, C.c_uint32, _vp, _vp]),
    "ofhe_ntt_inv": (C.c_int, [_vp, _vp, _vp, C.c_uint32, _vp, _vp]),
    "ofhe_ew_binop": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, C.c_uint32, _vp, _vp]),
    "ofhe_ew_mul_scalars": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint32, _vp, _vp]),
    "ofhe_behz_mtilde": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp, C.c_uint64, _vp, _vp]),
    "ofhe_behz_mtilde_fix": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ofhe_behz_sk_fix": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, C.c_uint64, _vp, _vp]),
    "ofhe_switch_modulus": (C.c_int, [_vp, _vp, _vp, C.c_uint32, C.c_uint32, _vp]),
    "ofhe_automorphism": (C.c_int, [_vp, _vp, _vp, C.c_uint32, C.c_uint32, _vp]),
    "ofhe_automorphism_coeff": (C.c_int, [_vp, _vp, _vp, C.c_uint32, _vp, C.c_uint32, _vp]),
    "ofhe_automorphism_map_host": (C.c_int, [_vp, C.c_uint32, _vp]),
    "ofhe_approx_switch_crt_basis": (C.c_int, [_vp, C.c_uint32, _vp, C.c_uint32, _vp, _vp, _vp, _vp]),
    "ofhe_bconv_create": (C.c_int, [_vp, C.c_uint32, _vp, _vp, C.c_uint32, _vp, _vp, C.c_int, _vp, _vp, C.POINTER(_vp)]),
    "ofhe_scale_round_create": (C.c_int, [_vp, C.c_uint32, _vp, C.c_uint32, _vp, C.c_int, _vp, _vp, C.POINTER(_vp)]),
    "ofhe_bconv_apply": (C.c_int, [_vp, _vp, _vp, _vp]),
    "ofhe_bconv_destroy": (None, [_vp]),
    "ofhe_hks_create": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(_vp)]),
    "ofhe_hks_destroy": (None, [_vp]),
    "ofhe_hks_num_per_part": (C.c_uint32, [_vp]),
    "ofhe_hks_num_digits": (C.c_uint32, [_vp, C.c_uint32]),
    "ofhe_hks_table_host": (C.c_int, [_vp, C.c_int, C.c_uint32, C.c_uint32, _vp, C.POINTER(C.c_size_t)]),
    "ofhe_evk_create": (C.c_int, [_vp, _vp, _vp, C.POINTER(_vp)]),
    "ofhe_evk_create_random": (C.c_int, [_vp, C.c_uint64, C.POINTER(_vp)]),
    "ofhe_evk_destroy": (None, [_vp]),
    "ofhe_hks_reserve": (C.c_int, [_vp, C.c_uint32, _vp]),
    "ofhe_evk_copy_to_device": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "ofhe_ntt_cluster_occupancy": (C.c_int, [_vp, C.c_uint32, _vp]),
    "ofhe_hks_modup": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp]),
    "ofhe_hks_inner_product": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp]),
    "ofhe_hks_approx_mod_down": (C.c_int, [_vp, C.c_uint32, _vp, _vp, C.c_uint32, _vp]),
    "ofhe_hks_keyswitch_core": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp]),
    "ofhe_ckks_evalmult": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp]),
    "ofhe_ckks_evalmult_host": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp]),
    "ofhe_ckks_evalmult_host_batch": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp, _vp]),
    "ofhe_ckks_rescale": (C.c_int, [_vp, C.c_uint32, _vp, _vp, C.c_uint32, _vp]),
    "ofhe_ckks_rotate": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp, _vp]),
    "ofhe_ckks_fast_rotation": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp, _vp, _vp]),
    "ofhe_ckks_fast_rotation_ext": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "ofhe_ckks_bsgs_workspace_words": (C.c_size_t, [_vp, C.c_uint32, C.c_uint32]),
    "ofhe_ckks_bsgs_level": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp, _vp, _vp, _vp, _vp, C.c_uint32, _vp,
                                       _vp, _vp, _vp]),
    "ofhe_hks_keyswitch_ext": (C.c_int, [_vp, C.c_uint32, C.c_int, _vp, _vp, _vp]),
    "ofhe_find_automorphism_index": (C.c_uint32, [C.c_int32, C.c_uint32]),
    "ofhe_ckks_paramgen": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, _vp, _vp,
                                      _vp]),
    "ofhe_root_of_unity": (C.c_uint64, [C.c_uint64, C.c_uint64]),
    "ofhe_first_prime": (C.c_uint64, [C.c_uint32, C.c_uint64]),
    "ofhe_last_prime": (C.c_uint64, [C.c_uint32, C.c_uint64]),
    "ofhe_previous_prime": (C.c_uint64, [C.c_uint64, C.c_uint64]),
    "ofhe_next_prime": (C.c_uint64, [C.c_uint64, C.c_uint64]),
    "ofhe_shard_create": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(_vp)]),
    "ofhe_shard_destroy": (None, [_vp]),
    "ofhe_shard_info": (C.c_int, [_vp, C.c_uint32, _vp]),
    "ofhe_shard_owned": (C.c_int, [_vp, C.c_uint32, _vp, _vp]),
    "ofhe_shard_evalmult_a": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp]),
    "ofhe_shard_evalmult_b": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp]),
    "ofhe_shard_evalmult_c": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp]),
    "ofhe_cereal_dcrtpoly_scan": (C.c_int, [_vp, C.c_size_t, C.c_uint32, _vp, _vp, _vp, _vp]),
    "ofhe_cereal_dcrtpoly_write": (C.c_int, [C.c_uint32, C.c_uint32, _vp, _vp, C.c_uint32, _vp, _vp, C.c_size_t,
                                             C.POINTER(C.c_size_t)]),
    "ofhe_dcrtpoly_upload_cereal": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.c_uint32, C.c_int, _vp, _vp, _vp]),
    "ofhe_bfv_tables_host": (C.c_int, [C.c_uint32, _vp, C.c_uint32, _vp, C.c_uint64, C.c_int, _vp, C.POINTER(C.c_size_t)]),
    "ofhe_bfv_create": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    "ofhe_bfv_destroy": (None, [_vp]),
    "ofhe_bfv_evalmult_norelin": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "ofhe_bfv_evalmult": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ofhe_shard_evalmult_a_push": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp]),
    "ofhe_shard_evalmult_b_push": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp]),
    "ofhe_shard_rotate_a": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp]),
    "ofhe_shard_rotate_a_push": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp]),
    "ofhe_shard_rotate_c": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp, _vp]),
    "ofhe_shard_rescale_a": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp]),
    "ofhe_shard_rescale_a_push": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp]),
    "ofhe_shard_rescale_b": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp]),
    "ofhe_peer_flags_words": (C.c_size_t, [C.c_uint32, C.c_uint32]),
    "ofhe_peer_barrier": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp]),
    "ofhe_peer_barrier_timeouts": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, _u64p]),
    "ofhe_ipc_export": (C.c_int, [_vp, _vp, _vp]),
    "ofhe_ipc_import": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "ofhe_ipc_release": (C.c_int, [_vp, _vp]),
    "ofhe_fill_uniform": (C.c_int, [_vp, _vp, C.c_uint32, _vp, C.c_uint64, _vp]),
}

_lib = None


class OfheError(RuntimeError):
    pass


def lib():
    """Load libofhe_b200.so (built in-tree by `make -C openfhe-development_b200`); fails loudly if missing."""
    global _lib
    if _lib is None:
        path = os.environ.get("OFHE_B200_LIB", LIB_PATH)  # measurement builds (tools/variants.sh); default: in-tree
        if not os.path.exists(path):
            raise OfheError(f"{path} is missing: build it with __graft_entry__.build() "
                            "(there is no CPU fallback for this backend)")
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _ck(rc):
    if rc != 0:
        raise OfheError(lib().ofhe_last_error().decode())


def _ptr(x):
    """Device pointer of a torch tensor / raw int; host pointer of a numpy array."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()  # torch tensor


def _idx(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def ptr_array(ptrs):
    """Host array of device pointers (the `*_of_rank_h` arguments)."""
    return (C.c_void_p * len(ptrs))(*[int(_ptr(p)) for p in ptrs])


def _stream():
    import torch

    return torch.cuda.current_stream().cuda_stream


# ---- host helpers (setup time, no GPU needed) ---------------------------------------------------------------------
def ckks_paramgen(N, depth, scale_bits, first_bits, dnum=0, scal_tech=0):
    """CKKS modulus chain with the reference's selection rules -> dict(moduli, roots, sizeQ, sizeP, dnum, alpha)."""
    moduli = np.zeros(depth + 1 + 64, dtype=np.uint64)
    roots = np.zeros_like(moduli)
    shape = np.zeros(4, dtype=np.uint32)
    _ck(lib().ofhe_ckks_paramgen(N, depth, scale_bits, first_bits, dnum, scal_tech, moduli.ctypes.data,
                                 roots.ctypes.data, shape.ctypes.data))
    sizeQ, sizeP, dn, alpha = (int(v) for v in shape)
    n = sizeQ + sizeP
    return dict(N=N, moduli=moduli[:n].copy(), roots=roots[:n].copy(), sizeQ=sizeQ, sizeP=sizeP, dnum=dn, alpha=alpha)


def dcrt_params(N, n_limbs, bits):
    """ILDCRTParams(order, depth, bits) moduli: LastPrime(bits, 2N) then successive PreviousPrime
    (lattice/hal/default/ildcrtparams.h:100-118), with their minimal roots."""
    L = lib()
    m = 2 * N
    q = [L.ofhe_last_prime(bits, m)]
    for _ in range(n_limbs - 1):
        q.append(L.ofhe_previous_prime(q[-1], m))
    roots = [L.ofhe_root_of_unity(m, x) for x in q]
    return np.array(q, dtype=np.uint64), np.array(roots, dtype=np.uint64)


def find_automorphism_index(index, m):
    return lib().ofhe_find_automorphism_index(index, m)


# ---- device objects -------------------------------------------------------------------------------------------------
class Context:
    def __init__(self, device=0):
        h = _vp()
        _ck(lib().ofhe_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device

    @property
    def sm_count(self):
        return lib().ofhe_ctx_sm_count(self.h)

    def close(self):
        if self.h:
            lib().ofhe_ctx_destroy(self.h)
            self.h = None

    def zeros(self, *shape):
        """uint64 slab on this device (torch int64 storage, bit pattern = uint64)."""
        import torch

        return torch.zeros(*shape, dtype=torch.int64, device=f"cuda:{self.device}")

    def upload(self, arr):
        import torch

        a = np.ascontiguousarray(arr, dtype=np.uint64)
        return torch.from_numpy(a.view(np.int64)).to(f"cuda:{self.device}")

    @staticmethod
    def download(t):
        return t.detach().cpu().numpy().view(np.uint64)

    # raw cudaMalloc buffers (zero-filled) for the peer-push exchange: IPC handles exist only for whole allocations
    def raw_zeros(self, words):
        p = _vp()
        _ck(lib().ofhe_malloc(self.h, 8 * int(words), C.byref(p)))
        z = np.zeros(int(words), dtype=np.uint64)
        _ck(lib().ofhe_memcpy_h2d(self.h, p, z.ctypes.data, z.nbytes, None))
        _ck(lib().ofhe_stream_sync(self.h, None))
        return p.value

    def raw_free(self, ptr):
        _ck(lib().ofhe_free(self.h, ptr))

    def raw_download(self, ptr, words):
        o = np.empty(int(words), dtype=np.uint64)
        _ck(lib().ofhe_memcpy_d2h(self.h, o.ctypes.data, ptr, o.nbytes, None))
        _ck(lib().ofhe_stream_sync(self.h, None))
        return o

    def ipc_export(self, ptr):
        h = C.create_string_buffer(64)
        _ck(lib().ofhe_ipc_export(self.h, ptr, h))
        return h.raw

    def ipc_import(self, handle):
        p = _vp()
        _ck(lib().ofhe_ipc_import(self.h, C.create_string_buffer(handle, 64), C.byref(p)))
        return p.value

    def ipc_release(self, ptr):
        _ck(lib().ofhe_ipc_release(self.h, ptr))

    def peer_barrier(self, flags_of_rank, rank, index):
        """flags_of_rank: ctypes array from ptr_array()."""
        _ck(lib().ofhe_peer_barrier(self.h, flags_of_rank, len(flags_of_rank), rank, index, _stream()))

    def peer_barrier_timeouts(self, flags_ptr, world, n_barriers):
        n = C.c_uint64(0)
        _ck(lib().ofhe_peer_barrier_timeouts(self.h, flags_ptr, world, n_barriers, C.byref(n)))
        return n.value


class Ring:
    """ILDCRTParams + twiddle tables on the device."""

    def __init__(self, ctx, N, moduli, roots):
        self.ctx = ctx
        self.N = int(N)
        self.moduli = np.ascontiguousarray(moduli, dtype=np.uint64)
        self.roots = np.ascontiguousarray(roots, dtype=np.uint64)
        h = _vp()
        _ck(lib().ofhe_ring_create(ctx.h, self.N, len(self.moduli), self.moduli.ctypes.data, self.roots.ctypes.data,
                                   C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            lib().ofhe_ring_destroy(self.h)
            self.h = None

    def twiddles(self, mod_idx):
        T, Tp, Ti, Tip = (np.empty(self.N, dtype=np.uint64) for _ in range(4))
        ni, nip = C.c_uint64(), C.c_uint64()
        _ck(lib().ofhe_ring_twiddles_host(self.h, mod_idx, T.ctypes.data, Tp.ctypes.data, Ti.ctypes.data,
                                          Tip.ctypes.data, C.byref(ni), C.byref(nip)))
        return T, Tp, Ti, Tip, ni.value, nip.value

    def upload_cereal(self, buf, mod_idx=None, check_residues=True):
        """Serialized DCRTPoly (the reference's portable-binary archive) -> [n_towers][N] device slab of THIS ring.
        The archive is rejected unless its ring dimension and tower moduli match the ring's (mod_idx: ring index of
        every tower, default 0..n-1).  Pass a pinned uint8 torch tensor for truly asynchronous copies."""
        import torch

        b = buf.numpy() if hasattr(buf, "numpy") else np.frombuffer(buf, dtype=np.uint8)
        n = len(self.moduli) if mod_idx is None else len(mod_idx)
        dst = self.ctx.zeros(n, self.N)
        shape = (C.c_uint32 * 3)()
        mi = None if mod_idx is None else _idx(mod_idx)
        _ck(lib().ofhe_dcrtpoly_upload_cereal(self.h, None if mi is None else mi.ctypes.data, b.ctypes.data, b.size, n,
                                              1 if check_residues else 0, _ptr(dst), shape, _stream()))
        torch.cuda.current_stream().synchronize()  # `b` may be a temporary view of the caller's bytes
        return dst[: shape[1]], {"N": shape[0], "n_towers": shape[1], "format": shape[2]}

    # DCRTPoly::SwitchFormat pieces
    def ntt_fwd(self, src, dst, mod_idx):
        mi = _idx(mod_idx)
        _ck(lib().ofhe_ntt_fwd(self.h, _ptr(src), _ptr(dst), len(mi), mi.ctypes.data, _stream()))

    def ntt_inv(self, src, dst, mod_idx):
        mi = _idx(mod_idx)
        _ck(lib().ofhe_ntt_inv(self.h, _ptr(src), _ptr(dst), len(mi), mi.ctypes.data, _stream()))

    def binop(self, op, dst, a, b, mod_idx):
        mi = _idx(mod_idx)
        _ck(lib().ofhe_ew_binop(self.h, op, _ptr(dst), _ptr(a), _ptr(b), len(mi), mi.ctypes.data, _stream()))

    def mul_scalars(self, dst, a, scalars, mod_idx):
        mi = _idx(mod_idx)
        s = np.ascontiguousarray(scalars, dtype=np.uint64)
        _ck(lib().ofhe_ew_mul_scalars(self.h, _ptr(dst), _ptr(a), s.ctypes.data, len(mi), mi.ctypes.data, _stream()))

    def switch_modulus(self, dst, src, from_idx, to_idx):
        _ck(lib().ofhe_switch_modulus(self.h, _ptr(dst), _ptr(src), from_idx, to_idx, _stream()))

    def automorphism(self, dst, src, n_limbs, k):
        _ck(lib().ofhe_automorphism(self.h, _ptr(dst), _ptr(src), n_limbs, k, _stream()))

    def automorphism_coeff(self, dst, src, mod_idx, k):
        mi = _idx(mod_idx)
        _ck(lib().ofhe_automorphism_coeff(self.h, _ptr(dst), _ptr(src), len(mi), mi.ctypes.data, k, _stream()))

    def automorphism_map(self, k):
        m = np.empty(self.N, dtype=np.uint32)
        _ck(lib().ofhe_automorphism_map_host(self.h, k, m.ctypes.data))
        return m

    def approx_switch_crt_basis(self, src_idx, dst_idx, src, dst):
        si, di = _idx(src_idx), _idx(dst_idx)
        _ck(lib().ofhe_approx_switch_crt_basis(self.h, len(si), si.ctypes.data, len(di), di.ctypes.data, _ptr(src),
                                               _ptr(dst), _stream()))

    def fill_uniform(self, dst, mod_idx, seed):
        mi = _idx(mod_idx)
        _ck(lib().ofhe_fill_uniform(self.h, _ptr(dst), len(mi), mi.ctypes.data, seed, _stream()))


class BaseConv:
    """Table-driven base conversion (a10 with explicit tables, a12 exact SwitchCRTBasis, a13 ScaleAndRound)."""

    def __init__(self, h):
        self.h = h

    @classmethod
    def switch_crt_basis(cls, ring, src_idx, QHatInvModq, dst_idx, QHatModp, dst_major, alphaQModp=None, qInv=None):
        si, di = _idx(src_idx), _idx(dst_idx)
        hi = np.ascontiguousarray(QHatInvModq, dtype=np.uint64)
        m = np.ascontiguousarray(QHatModp, dtype=np.uint64)
        a = None if alphaQModp is None else np.ascontiguousarray(alphaQModp, dtype=np.uint64)
        d = None if qInv is None else np.ascontiguousarray(qInv, dtype=np.float64)
        h = _vp()
        _ck(lib().ofhe_bconv_create(ring.h, len(si), si.ctypes.data, hi.ctypes.data, len(di), di.ctypes.data,
                                    m.ctypes.data, int(dst_major), None if a is None else a.ctypes.data,
                                    None if d is None else d.ctypes.data, C.byref(h)))
        return cls(h)

    @classmethod
    def scale_and_round(cls, ring, in_idx, out_idx, out_limbs_first, tbl, frac):
        ii, oi = _idx(in_idx), _idx(out_idx)
        t = np.ascontiguousarray(tbl, dtype=np.uint64)
        f = np.ascontiguousarray(frac, dtype=np.float64)
        h = _vp()
        _ck(lib().ofhe_scale_round_create(ring.h, len(ii), ii.ctypes.data, len(oi), oi.ctypes.data,
                                          int(out_limbs_first), t.ctypes.data, f.ctypes.data, C.byref(h)))
        return cls(h)

    def apply(self, src, dst):
        _ck(lib().ofhe_bconv_apply(self.h, _ptr(src), _ptr(dst), _stream()))

    def close(self):
        if self.h:
            lib().ofhe_bconv_destroy(self.h)
            self.h = None


def cereal_scan_dcrtpoly(buf, max_towers=256):
    """Serialized DCRTPoly (bytes-like) -> dict(N, n_towers, format, moduli, roots, offsets): host parsing only."""
    b = np.frombuffer(buf, dtype=np.uint8)
    shape = np.zeros(3, dtype=np.uint32)
    q, r = np.zeros(max_towers, dtype=np.uint64), np.zeros(max_towers, dtype=np.uint64)
    off = np.zeros(max_towers, dtype=np.uintp)
    _ck(lib().ofhe_cereal_dcrtpoly_scan(b.ctypes.data, b.size, max_towers, shape.ctypes.data, q.ctypes.data,
                                        r.ctypes.data, off.ctypes.data))
    n = int(shape[1])
    return dict(N=int(shape[0]), n_towers=n, format=int(shape[2]), moduli=q[:n].copy(), roots=r[:n].copy(),
                offsets=off[:n].astype(np.int64))


def cereal_write_dcrtpoly(moduli, roots, fmt, data):
    """[n_towers][N] uint64 residues -> the reference's portable-binary archive (bytes)."""
    q = np.ascontiguousarray(moduli, dtype=np.uint64)
    r = np.ascontiguousarray(roots, dtype=np.uint64)
    d = np.ascontiguousarray(data, dtype=np.uint64)
    L, N = d.shape
    out = np.empty(L * (8 * N + 128) + 64 * L + 256, dtype=np.uint8)
    n = C.c_size_t(0)
    _ck(lib().ofhe_cereal_dcrtpoly_write(N, L, q.ctypes.data, r.ctypes.data, int(fmt), d.ctypes.data, out.ctypes.data,
                                         out.size, C.byref(n)))
    return out[:n.value].tobytes()


BFV_TABLE_NAMES = ["QlHatInvModq", "QlHatModr", "alphaQlModr", "qInv", "mNegRlQlHatInvModq", "qInvModr", "RlHatInvModr",
                   "RlHatModq", "alphaRlModq", "rInv", "tQlSlHatInvModsDivsModq", "tQlSlHatInvModsDivsFrac"]


def bfv_tables(moduliQ, moduliR, t):
    """HPSPOVERQ multiplication tables at the full level from (Q, R, t) alone (host arithmetic, no GPU needed)."""
    Q = np.ascontiguousarray(moduliQ, dtype=np.uint64)
    R = np.ascontiguousarray(moduliR, dtype=np.uint64)
    nQ, nR = len(Q), len(R)
    shapes = [(nQ,), (nR, nQ), (nQ + 1, nR), (nQ,), (nQ,), (nQ, nR), (nR,), (nQ, nR), (nR + 1, nQ), (nR,), (nQ, nR + 1),
              (nR,)]
    out = {}
    for w, (name, shp) in enumerate(zip(BFV_TABLE_NAMES, shapes)):
        a = np.empty(shp, dtype=np.float64 if w in (3, 9, 11) else np.uint64)
        n = C.c_size_t(0)
        _ck(lib().ofhe_bfv_tables_host(nQ, Q.ctypes.data, nR, R.ctypes.data, int(t), w, a.ctypes.data, C.byref(n)))
        assert n.value == a.size
        out[name] = a
    return out


class BfvMult:
    """BFV EvalMult, HPSPOVERQ (LeveledSHEBFVRNS::EvalMult): `ring` registers Q then R; `T` holds the reference's
    CryptoParametersBFVRNS tables at level sizeQ-1 under the getter names without the `Get` prefix."""

    def __init__(self, ring, sizeQ, sizeR, T):
        self.ring, self.sizeQ, self.sizeR = ring, sizeQ, sizeR
        iq, ir = list(range(sizeQ)), list(range(sizeQ, sizeQ + sizeR))
        B = BaseConv
        self.convs = [
            B.switch_crt_basis(ring, iq, T["QlHatInvModq"], ir, T["QlHatModr"], True, T["alphaQlModr"], T["qInv"]),
            B.switch_crt_basis(ring, iq, T["mNegRlQlHatInvModq"], ir, T["qInvModr"], False),
            B.switch_crt_basis(ring, ir, T["RlHatInvModr"], iq, T["RlHatModq"], True, T["alphaRlModq"], T["rInv"]),
            B.scale_and_round(ring, ir, iq, True, T["tQlSlHatInvModsDivsModq"], T["tQlSlHatInvModsDivsFrac"]),
        ]
        h = _vp()
        _ck(lib().ofhe_bfv_create(ring.h, sizeQ, sizeR, *[c.h for c in self.convs], C.byref(h)))
        self.h = h

    @classmethod
    def from_moduli(cls, ring, sizeQ, sizeR, t):
        """Tables derived by the library itself from the ring's moduli (Q then R) and the plaintext modulus t."""
        return cls(ring, sizeQ, sizeR, bfv_tables(ring.moduli[:sizeQ], ring.moduli[sizeQ:sizeQ + sizeR], t))

    def close(self):
        if self.h:
            lib().ofhe_bfv_destroy(self.h)
            self.h = None
        for c in self.convs:
            c.close()
        self.convs = []

    def evalmult_norelin(self, ct1, ct2, out3):
        _ck(lib().ofhe_bfv_evalmult_norelin(self.h, _ptr(ct1), _ptr(ct2), _ptr(out3), _stream()))

    def evalmult(self, hks, ct1, ct2, evk, out):
        _ck(lib().ofhe_bfv_evalmult(self.h, hks.h, _ptr(ct1), _ptr(ct2), evk.h, _ptr(out), _stream()))


class EvalKey:
    def __init__(self, hks, a=None, b=None, seed=None):
        h = _vp()
        if a is not None:
            a = np.ascontiguousarray(a, dtype=np.uint64)
            b = np.ascontiguousarray(b, dtype=np.uint64)
            _ck(lib().ofhe_evk_create(hks.h, a.ctypes.data, b.ctypes.data, C.byref(h)))
        else:
            _ck(lib().ofhe_evk_create_random(hks.h, 0 if seed is None else seed, C.byref(h)))
        self.h = h

    def copy_to(self, which, dst):
        """A (which=0) / B (which=1) vector -> dst [dnum][sizeQ+sizeP][N] (device tensor)."""
        _ck(lib().ofhe_evk_copy_to_device(self.h, which, _ptr(dst), _stream()))

    def close(self):
        if self.h:
            lib().ofhe_evk_destroy(self.h)
            self.h = None


class Hks:
    """KeySwitchHYBRID + CryptoParametersRNS tables over a ring whose moduli are [Q | P]."""

    def __init__(self, ring, sizeQ, sizeP, dnum):
        self.ring = ring
        self.sizeQ, self.sizeP, self.dnum = int(sizeQ), int(sizeP), int(dnum)
        h = _vp()
        _ck(lib().ofhe_hks_create(ring.h, sizeQ, sizeP, dnum, C.byref(h)))
        self.h = h
        self.alpha = lib().ofhe_hks_num_per_part(h)

    def close(self):
        if self.h:
            lib().ofhe_hks_destroy(self.h)
            self.h = None

    def reserve(self, sizeQl):
        """Level tables + this stream's workspace ahead of time (needed before capturing into a CUDA graph)."""
        _ck(lib().ofhe_hks_reserve(self.h, sizeQl, _stream()))

    def num_digits(self, sizeQl):
        return lib().ofhe_hks_num_digits(self.h, sizeQl)

    def table(self, which, a=0, b=0):
        n = C.c_size_t()
        _ck(lib().ofhe_hks_table_host(self.h, which, a, b, None, C.byref(n)))
        out = np.empty(n.value, dtype=np.uint64)
        _ck(lib().ofhe_hks_table_host(self.h, which, a, b, out.ctypes.data, C.byref(n)))
        return out

    def modup(self, sizeQl, c, digits):
        _ck(lib().ofhe_hks_modup(self.h, sizeQl, _ptr(c), _ptr(digits), _stream()))

    def inner_product(self, sizeQl, digits, evk, e01):
        _ck(lib().ofhe_hks_inner_product(self.h, sizeQl, _ptr(digits), evk.h, _ptr(e01), _stream()))

    def approx_mod_down(self, sizeQl, src, dst, n_polys=1):
        _ck(lib().ofhe_hks_approx_mod_down(self.h, sizeQl, _ptr(src), _ptr(dst), n_polys, _stream()))

    def keyswitch_core(self, sizeQl, c, evk, out01):
        _ck(lib().ofhe_hks_keyswitch_core(self.h, sizeQl, _ptr(c), evk.h, _ptr(out01), _stream()))

    def evalmult(self, sizeQl, ct1, ct2, evk, out):
        _ck(lib().ofhe_ckks_evalmult(self.h, sizeQl, _ptr(ct1), _ptr(ct2), evk.h, _ptr(out), _stream()))

    def evalmult_host(self, sizeQl, ct1, ct2, evk, out):
        """numpy (or pinned torch CPU) buffers in, numpy out: host<->device copies included."""
        _ck(lib().ofhe_ckks_evalmult_host(self.h, sizeQl, _ptr(ct1), _ptr(ct2), evk.h, _ptr(out)))

    def evalmult_host_batch(self, sizeQl, ct1s, ct2s, evk, outs):
        """n independent multiplications on (pinned) host buffers, H2D / compute / D2H software-pipelined."""
        n = len(ct1s)
        arr = lambda xs: (C.c_void_p * n)(*[_ptr(x) for x in xs])
        _ck(lib().ofhe_ckks_evalmult_host_batch(self.h, sizeQl, n, arr(ct1s), arr(ct2s), evk.h, arr(outs)))

    def rescale(self, sizeQl, src, dst, n_polys=2):
        _ck(lib().ofhe_ckks_rescale(self.h, sizeQl, _ptr(src), _ptr(dst), n_polys, _stream()))

    def rotate(self, sizeQl, k, ct, evk, out):
        _ck(lib().ofhe_ckks_rotate(self.h, sizeQl, k, _ptr(ct), evk.h, _ptr(out), _stream()))

    # hoisted rotations: digits = modup(ct[1]) once (EvalFastRotationPrecompute), then any number of rotations
    def fast_rotation(self, sizeQl, k, ct, digits, evk, out):
        _ck(lib().ofhe_ckks_fast_rotation(self.h, sizeQl, k, _ptr(ct), _ptr(digits), evk.h, _ptr(out), _stream()))

    def fast_rotation_ext(self, sizeQl, k, add_first, ct, digits, evk, out_ext):
        _ck(lib().ofhe_ckks_fast_rotation_ext(self.h, sizeQl, k, int(add_first), _ptr(ct), _ptr(digits), evk.h,
                                              _ptr(out_ext), _stream()))

    def bsgs_workspace(self, sizeQl, g):
        """device workspace for bsgs_level (caller-owned, reusable)"""
        words = lib().ofhe_ckks_bsgs_workspace_words(self.h, sizeQl, g)
        return self.ring.ctx.zeros(words)

    def bsgs_level(self, sizeQl, k_in, k_out, evk_in, evk_out, A, skip, ct, out, ws):
        """One level of the double-hoisted BSGS linear transform (ckksrns-fhe.cpp:1934-1987), device resident.
        k_*: automorphism indices (0 = no rotation); evk_*: EvalKey or None per step; A [b*g][sizeQl+sizeP][N]."""
        g, b = len(k_in), len(k_out)
        ki, ko = _idx(k_in), _idx(k_out)
        pin = (_vp * g)(*[(e.h if e is not None else None) for e in evk_in])
        pout = (_vp * b)(*[(e.h if e is not None else None) for e in evk_out])
        _ck(lib().ofhe_ckks_bsgs_level(self.h, sizeQl, g, b, ki.ctypes.data, ko.ctypes.data, pin, pout, _ptr(A), skip,
                                       _ptr(ct), _ptr(out), _ptr(ws), _stream()))

    def keyswitch_ext(self, sizeQl, add_first, ct, out_ext):
        _ck(lib().ofhe_hks_keyswitch_ext(self.h, sizeQl, int(add_first), _ptr(ct), _ptr(out_ext), _stream()))

    def keyswitch_down(self, sizeQl, ct_ext, out):
        self.approx_mod_down(sizeQl, ct_ext, out, n_polys=2)


# ---- limb-sharded EvalMult (SURVEY.md 8 e) --------------------------------------------------------------------------
def shard_plan(sizeQ, sizeP, sizeQl, world, rank):
    """Ownership of the limb-sharded layout (pure host logic, mirrored by csrc/shard.cu and checked against it):
    Q limb i -> rank i % world, local row i // world; P limb j -> rank (j + sizeQ) % world, local row j // world."""
    own_q = [i for i in range(sizeQl) if i % world == rank]
    own_q_full = [i for i in range(sizeQ) if i % world == rank]
    own_p = [j for j in range(sizeP) if (j + sizeQ) % world == rank]
    return dict(own_q=own_q, own_p=own_p, own_q_full=own_q_full, slots_q=-(-sizeQl // world), slots_p=-(-sizeP // world),
                n_key_rows=len(own_q_full) + len(own_p),
                gather_q_row=lambda i: (i % world) * (-(-sizeQl // world)) + i // world,
                gather_p_row=lambda j, b: (((j + sizeQ) % world) * 2 + b) * (-(-sizeP // world)) + j // world)


class PeerBuffers:
    """The same raw buffer on every rank, addressable from this process: the local cudaMalloc allocation plus CUDA-IPC
    mappings of the peers' (one process per GPU).  `exchange(obj) -> [obj of rank 0, ..., obj of rank world-1]` is the
    caller's control-plane all-gather (torch.distributed.all_gather_object)."""

    def __init__(self, ctx, words, rank, world, exchange):
        self.ctx, self.rank, self.world, self.words = ctx, rank, world, int(words)
        self.local = ctx.raw_zeros(max(self.words, 1))
        handles = exchange(ctx.ipc_export(self.local))
        self.ptrs = [self.local if r == rank else ctx.ipc_import(handles[r]) for r in range(world)]
        self.array = ptr_array(self.ptrs)

    def close(self):
        for r, p in enumerate(self.ptrs):
            if r != self.rank and p:
                self.ctx.ipc_release(p)
        if self.local:
            self.ctx.raw_free(self.local)
        self.ptrs, self.local = [], None


class Shard:
    """One rank's share of a limb-sharded CKKS EvalMult.  `all_gather(buf)` is supplied by the caller: in place on a
    [world, ...] tensor (torch.distributed.all_gather_into_tensor over NCCL in production; a local copy loop in the
    single-GPU virtual-rank tests)."""

    def __init__(self, ring, sizeQ, sizeP, dnum, rank, world):
        self.ring, self.sizeQ, self.sizeP, self.dnum, self.rank, self.world = ring, sizeQ, sizeP, dnum, rank, world
        h = _vp()
        _ck(lib().ofhe_shard_create(ring.h, sizeQ, sizeP, dnum, rank, world, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            lib().ofhe_shard_destroy(self.h)
            self.h = None

    def info(self, sizeQl):
        o = np.zeros(6, dtype=np.uint32)
        _ck(lib().ofhe_shard_info(self.h, sizeQl, o.ctypes.data))
        return dict(n_own_q=int(o[0]), n_own_p=int(o[1]), slots_q=int(o[2]), slots_p=int(o[3]), n_key_rows=int(o[4]),
                    n_digits=int(o[5]))

    def owned(self, sizeQl):
        i = self.info(sizeQl)
        q = np.zeros(max(i["n_own_q"], 1), dtype=np.uint32)
        p = np.zeros(max(i["n_own_p"], 1), dtype=np.uint32)
        _ck(lib().ofhe_shard_owned(self.h, sizeQl, q.ctypes.data, p.ctypes.data))
        return q[:i["n_own_q"]].tolist(), p[:i["n_own_p"]].tolist()

    def phase_a(self, sizeQl, ct1, ct2, out, gather_q):
        _ck(lib().ofhe_shard_evalmult_a(self.h, sizeQl, _ptr(ct1), _ptr(ct2), _ptr(out), _ptr(gather_q), _stream()))

    def phase_b(self, sizeQl, gather_q, evk_a, evk_b, gather_p):
        _ck(lib().ofhe_shard_evalmult_b(self.h, sizeQl, _ptr(gather_q), _ptr(evk_a), _ptr(evk_b), _ptr(gather_p),
                                        _stream()))

    def phase_c(self, sizeQl, gather_p, out):
        _ck(lib().ofhe_shard_evalmult_c(self.h, sizeQl, _ptr(gather_p), _ptr(out), _stream()))

    # ---- exchange by peer stores + device barrier (no collective call); gq/gp/flags: ptr_array() of every rank's buffer
    def phase_a_push(self, sizeQl, ct1, ct2, out, gq_of_rank):
        _ck(lib().ofhe_shard_evalmult_a_push(self.h, sizeQl, _ptr(ct1), _ptr(ct2), _ptr(out), gq_of_rank, _stream()))

    def phase_b_push(self, sizeQl, evk_a, evk_b, gq_of_rank, gp_of_rank):
        _ck(lib().ofhe_shard_evalmult_b_push(self.h, sizeQl, _ptr(evk_a), _ptr(evk_b), gq_of_rank, gp_of_rank,
                                             _stream()))

    def evalmult_push(self, sizeQl, ct1, ct2, evk_a, evk_b, out, gq_of_rank, gp_of_rank, flags_of_rank, slot=0):
        ctx = self.ring.ctx
        self.phase_a_push(sizeQl, ct1, ct2, out, gq_of_rank)
        ctx.peer_barrier(flags_of_rank, self.rank, 2 * slot)
        self.phase_b_push(sizeQl, evk_a, evk_b, gq_of_rank, gp_of_rank)
        ctx.peer_barrier(flags_of_rank, self.rank, 2 * slot + 1)
        self.phase_c(sizeQl, gp_of_rank[self.rank], out)

    # ---- limb-sharded rotation (EvalAutomorphism with automorphism index k) and rescale
    def rotate(self, sizeQl, k, ct, evk_a, evk_b, tmp, out, gather_q, gather_p, all_gather):
        L = lib()
        _ck(L.ofhe_shard_rotate_a(self.h, sizeQl, _ptr(ct), _ptr(tmp), _ptr(gather_q), _stream()))
        all_gather(gather_q)
        self.phase_b(sizeQl, gather_q, evk_a, evk_b, gather_p)
        all_gather(gather_p)
        _ck(L.ofhe_shard_rotate_c(self.h, sizeQl, k, _ptr(gather_p), _ptr(tmp), _ptr(out), _stream()))

    def rotate_push(self, sizeQl, k, ct, evk_a, evk_b, tmp, out, gq_of_rank, gp_of_rank, flags_of_rank, slot=0):
        ctx, L = self.ring.ctx, lib()
        _ck(L.ofhe_shard_rotate_a_push(self.h, sizeQl, _ptr(ct), _ptr(tmp), gq_of_rank, _stream()))
        ctx.peer_barrier(flags_of_rank, self.rank, 2 * slot)
        self.phase_b_push(sizeQl, evk_a, evk_b, gq_of_rank, gp_of_rank)
        ctx.peer_barrier(flags_of_rank, self.rank, 2 * slot + 1)
        _ck(L.ofhe_shard_rotate_c(self.h, sizeQl, k, gp_of_rank[self.rank], _ptr(tmp), _ptr(out), _stream()))

    def rescale(self, sizeQl, ct, out, gather_l, broadcast):
        """broadcast(buf, src_rank): the caller's collective (torch.distributed.broadcast in production)."""
        L = lib()
        _ck(L.ofhe_shard_rescale_a(self.h, sizeQl, _ptr(ct), _ptr(gather_l), _stream()))
        broadcast(gather_l, (sizeQl - 1) % self.world)
        _ck(L.ofhe_shard_rescale_b(self.h, sizeQl, _ptr(ct), _ptr(gather_l), _ptr(out), _stream()))

    def rescale_push(self, sizeQl, ct, out, gl_of_rank, flags_of_rank, index=0):
        ctx, L = self.ring.ctx, lib()
        _ck(L.ofhe_shard_rescale_a_push(self.h, sizeQl, _ptr(ct), gl_of_rank, _stream()))
        ctx.peer_barrier(flags_of_rank, self.rank, index)
        _ck(L.ofhe_shard_rescale_b(self.h, sizeQl, _ptr(ct), gl_of_rank[self.rank], _ptr(out), _stream()))

    def evalmult(self, sizeQl, ct1, ct2, evk_a, evk_b, out, gather_q, gather_p, all_gather):
        self.phase_a(sizeQl, ct1, ct2, out, gather_q)
        all_gather(gather_q)
        self.phase_b(sizeQl, gather_q, evk_a, evk_b, gather_p)
        all_gather(gather_p)
        self.phase_c(sizeQl, gather_p, out)
