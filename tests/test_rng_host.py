"""The device RNG header compiled for the host (it is __host__ __device__ code) against the oracle: the incremental
MT19937 twist and the paired draw reproduce the reference's batch generator for every alignment of the 624-word wrap."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O

HERE = Path(__file__).resolve().parent


@pytest.fixture(scope="module")
def drv(tmp_path_factory):
    so = tmp_path_factory.mktemp("rngdrv") / "librngdrv.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-o", str(so),
                           str(HERE / "rng_host_driver.cc")])
    L = C.CDLL(str(so))
    L.dev_mt_stream.argtypes = [C.c_uint32, C.c_int, C.c_uint32, C.c_uint64, C.c_void_p]
    L.dev_mtgen_stream.argtypes = [C.c_uint32, C.c_uint64, C.c_void_p]
    L.dev_mesh_routes.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p]
    L.dev_mesh_routes.restype = C.c_uint64
    L.dev_marsaglia_stream.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p]
    L.dev_xorshift_stream.argtypes = [C.c_uint32, C.c_uint64, C.c_void_p]
    return L


def test_incremental_twist_equals_batch_generator(drv):
    n = 624 * 7 + 13
    ref, _ = O.rng_u32_stream(0, 1447, 0, n)
    out = np.zeros(n, dtype=np.uint32)
    drv.dev_mt_stream(1447, 0, 0, n, out.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("lead", [0, 1, 2, 3, 226, 227, 228, 621, 622, 623, 624, 625])
def test_paired_draw_equals_two_single_draws(drv, lead):
    n = 624 * 5 + 7
    ref, _ = O.rng_u32_stream(0, 100 + lead, 0, n)
    out = np.zeros(n, dtype=np.uint32)
    drv.dev_mt_stream(100 + lead, 2, lead, n, out.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out, ref)


def test_whole_generation_storage_equals_reference_stream(drv):
    # rng.cuh MtGenDev: the reference's own representation (mersenne.cc:56-68), next generation built ahead of time
    n = 624 * 7 + 13
    ref, _ = O.rng_u32_stream(0, 4242, 0, n)
    out = np.zeros(n, dtype=np.uint32)
    drv.dev_mtgen_stream(4242, n, out.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("ng,lead", [(4, 4), (2, 2), (4, 0), (4, 422), (4, 424), (4, 622)])
def test_route_digest_windows_equal_sequential_draws(drv, ng, lead):
    # the event-parallel form of MessageMesh (window_parallel_kernel + maintenance_kernel) restated on the host: the
    # port of the j-th pair of draws must be (first draw of the pair) % ng whatever the window sizes are, across
    # generation boundaries, with windows run by the sequential form in between
    pairs = 312 * 9 + 5
    sizes = np.array([1, 7, 100, 3, 16, 17, 0, 99, 100, 100, 2, 15, 33, 64, 5, 100, 31], dtype=np.uint32)
    ref, _ = O.rng_u32_stream(0, 100 + lead + ng, 0, lead + 2 * pairs)
    out = np.zeros(pairs, dtype=np.uint32)
    got = drv.dev_mesh_routes(100 + lead + ng, lead, ng, pairs, sizes.ctypes.data_as(C.c_void_p), len(sizes),
                              out.ctypes.data_as(C.c_void_p))
    assert got == pairs
    assert np.array_equal(out, ref[lead:lead + 2 * pairs:2] % ng)


def test_marsaglia_and_xorshift(drv):
    n = 5000
    ref, _ = O.rng_u32_stream(1, 1053, 1447, n)
    out = np.zeros(n, dtype=np.uint32)
    uni = np.zeros(n, dtype=np.float64)
    drv.dev_marsaglia_stream(1053, 1447, n, out.ctypes.data_as(C.c_void_p), uni.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out, ref)
    ref, _ = O.rng_u32_stream(2, 1447, 0, n)
    drv.dev_xorshift_stream(1447, n, out.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out, ref)


def test_fastmod_equals_the_remainder_operator(drv):
    # PHOLD's `u32 % dmod` and the clock-driven elements' `i32 % commFreq == 0` use a multiply-high remainder
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.integers(0, 1 << 32, 200000, dtype=np.uint64).astype(np.uint32),
                         np.array([0, 1, 2, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF, 0xFFFFFFFE], dtype=np.uint32)])
    drv.dev_fastmod_check.argtypes = [C.c_uint32, C.c_void_p, C.c_uint64]
    for d in [1, 2, 3, 4, 5, 7, 1000, 8000, 65535, 65536, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF, 12345677]:
        assert drv.dev_fastmod_check(d, xs.ctypes.data_as(C.c_void_p), len(xs)) == 1, d
