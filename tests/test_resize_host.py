"""The arithmetic and indexing of resize_input.cu on the CPU: the marked region of the CUDA source (weight tables, clip8, the kernel body) is
compiled with g++ — the __d*_rn intrinsics mapped to plain IEEE operators, FMA contraction off, one sequential "thread" per CTA — and compared
with the oracle's tables (oracle/resize.py, itself pinned bit-exact against Pillow) and with the fixture recorded from the reference's own
resizeNormalize.  This is a check of the device SOURCE (it cannot see races or barrier mistakes — the GPU suite does); nothing here is shipped
or timed."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, 'tpgsr_b200', 'csrc', 'resize_input.cu')

PRELUDE = r'''
#include <math.h>
#include <string.h>
#include <stddef.h>
#define __global__
#define __launch_bounds__(x)
#define __shared__
#define __align__(x)
#define __syncthreads()
#define __shfl_xor_sync(m, v, o) (0)
#define __fdiv_rn(a, b) ((a) / (b))
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __dadd_rn(a, b) ((a) + (b))
#define __dmul_rn(a, b) ((a) * (b))
#define __ddiv_rn(a, b) ((a) / (b))
struct Dim { int x; };
static Dim threadIdx = {0}, blockDim = {1}, blockIdx = {0};
constexpr int PREC = 32 - 8 - 2;
constexpr int RS_THREADS = 256;
'''

WRAPPER = r'''
extern "C" int host_ksize(int in_size, int out_size) { return ksize_for(in_size, out_size); }
extern "C" void host_coeffs(int in_size, int out_size, int pitch, int* bounds, int* kk) { build_coeffs(in_size, out_size, pitch, bounds, kk); }
extern "C" int host_clip8(int v) { return clip8(v); }
unsigned char smem_rs[1 << 20];
extern "C" void host_resize(const unsigned char* src, const long long* offsets, const int* hs, const int* ws, int B, int max_h, int max_w, int H,
                            int W, int mask, float* out, unsigned char* tmp) {
    ResizeParams p;
    p.src = src; p.offsets = offsets; p.hs = hs; p.ws = ws; p.out = out; p.tmp = tmp;
    p.H = H; p.W = W; p.mask = mask; p.max_h = max_h; p.max_w = max_w;
    p.ksx = ksize_for(max_w, W); p.ksy = ksize_for(max_h, H);
    for (int b = 0; b < B; ++b) {
        blockIdx.x = b;
        memset(smem_rs, 0, sizeof(smem_rs));
        resize_normalize_kernel(p);
    }
}
'''


@pytest.fixture(scope='module')
def hostlib(tmp_path_factory):
    text = open(SRC).read()
    a, b = text.index('// [host-testable begin]'), text.index('// [host-testable end]')
    d = tmp_path_factory.mktemp('resize_host')
    cpp = d / 'resize_host.cpp'
    cpp.write_text(PRELUDE + text[a:b] + WRAPPER)
    so = d / 'resize_host.so'
    subprocess.check_call(['g++', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-o', str(so), str(cpp)])
    lib = ctypes.CDLL(str(so))
    lib.host_coeffs.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return lib


def test_device_weight_tables_equal_the_oracle(hostlib):
    from oracle import resize
    r = np.random.RandomState(11)
    cases = [(1, 64), (2, 64), (3, 16), (16, 16), (17, 16), (63, 64), (64, 128), (127, 128), (129, 128), (300, 64), (299, 128), (75, 32),
             (1000, 128), (3000, 128), (5, 32), (31, 32), (33, 32)]
    cases += [(int(r.randint(1, 400)), int(o)) for o in (16, 32, 64, 128) for _ in range(25)]
    for in_size, out_size in cases:
        ks, bounds, kk = resize.coeffs(in_size, out_size)
        assert hostlib.host_ksize(in_size, out_size) == ks, (in_size, out_size)
        pitch = ks + 3                                       # a larger pitch (tables are sized for the largest crop of the batch)
        hb = np.full((out_size, 2), -1, dtype=np.int32)
        hk = np.zeros((out_size, pitch), dtype=np.int32)
        hostlib.host_coeffs(in_size, out_size, pitch, hb.ctypes.data, hk.ctypes.data)
        assert np.array_equal(hb, bounds), (in_size, out_size)
        for xx in range(out_size):
            n = bounds[xx, 1]
            assert np.array_equal(hk[xx, :n], kk[xx, :n]), (in_size, out_size, xx)


def test_device_clip8(hostlib):
    for v in (-(1 << 30), -(1 << 22) - 1, -1, 0, (1 << 22) - 1, 1 << 22, 255 << 22, (256 << 22) - 1, 256 << 22, (1 << 31) - 1):
        want = min(max(v >> 22, 0), 255)
        assert hostlib.host_clip8(v) == want, v


def _run_host(hostlib, crops, size, mask, max_hw=None):
    W, H = size
    B = len(crops)
    hs = np.array([c.shape[0] for c in crops], dtype=np.int32)
    ws = np.array([c.shape[1] for c in crops], dtype=np.int32)
    sizes = hs.astype(np.int64) * ws * 3
    offsets = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    src = np.concatenate([c.reshape(-1) for c in crops]).astype(np.uint8)
    max_h, max_w = max_hw or (int(hs.max()), int(ws.max()))
    out = np.full((B, 4 if mask else 3, H, W), -5.0, dtype=np.float32)
    tmp = np.zeros(B * max_h * W * 3, dtype=np.uint8)
    hostlib.host_resize.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 6 + [ctypes.c_void_p] * 2
    hostlib.host_resize(src.ctypes.data, offsets.ctypes.data, hs.ctypes.data, ws.ctypes.data, B, max_h, max_w, H, W, int(mask), out.ctypes.data,
                        tmp.ctypes.data)
    return out


def test_device_kernel_body_reproduces_the_reference_fixture(hostlib):
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'resize_normalize.npz'))
    crops = [g[f'x/{i}'] for i in range(int(g['n']))]
    for tag, size in (('hr', (128, 32)), ('lr', (64, 16))):
        for mask in (0, 1):
            want = g[f'y/{tag}/{mask}'].astype(np.float32) / np.float32(255)
            got = _run_host(hostlib, crops, size, mask)
            assert np.array_equal(got, want), (tag, mask, [i for i in range(len(crops)) if not np.array_equal(got[i], want[i])])


def test_device_kernel_body_poisons_unprovisioned_sizes(hostlib):
    crops = [np.full((4, 9, 3), 7, dtype=np.uint8), np.full((3, 5, 3), 9, dtype=np.uint8)]
    out = _run_host(hostlib, crops, (8, 4), 1, max_hw=(3, 9))        # crop 0 is taller than the launch was provisioned for
    assert np.isnan(out[0]).all() and np.isfinite(out[1]).all()
