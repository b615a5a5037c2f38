"""ThreadSanitizer over the GPU-validated kernels that the CPU tier executes under tests/cpu_emul/cuda_shim.h (one OS thread per CUDA
thread): a `__syncthreads()` / `__syncwarp()` missing between a write and another thread's read would be a data race between the
emulating threads.  (The persistent ECO kernels have their own sanitizer runs in tests/test_eco_cpu.py, incl. the grid barriers.)"""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESSES = ("transformer", "corr", "prroi", "atom", "conv_fp32", "sd", "cg", "eco_loc")


def _build_all(tmp_path, flags):
    from concurrent.futures import ThreadPoolExecutor

    def one(n):
        subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas"] + flags +
                       [os.path.join(ROOT, "tests", "cpu_emul", n + "_emul.cpp"), "-o", str(tmp_path / ("lib%s_tsan.so" % n))],
                       check=True, capture_output=True)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        list(ex.map(one, HARNESSES))


def test_block_level_kernels_under_thread_sanitizer(tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    rt = subprocess.run(["g++", "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(rt) or not os.path.exists(rt):
        pytest.skip("no ThreadSanitizer runtime")
    _build_all(tmp_path, ["-fsanitize=thread,alignment", "-Wno-tsan"])
    # B200_EMUL_THREADS: real OS threads per block (the default cooperative fiber mode of launch_blocks has nothing for the sanitizer to see)
    env = dict(os.environ, LD_PRELOAD=rt, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0", B200_EMUL_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "cpu_emul", "tsan_sweep.py"), str(tmp_path)], capture_output=True, text=True,
                       env=env, timeout=900)
    if "EMUL_DONE" not in r.stdout and "ThreadSanitizer" not in r.stderr:
        pytest.skip("ThreadSanitizer could not run here: %s" % r.stderr[-300:])
    assert "EMUL_DONE" in r.stdout, r.stderr[-2000:]
    assert "data race" not in r.stderr, r.stderr[:4000]
    assert "runtime error" not in r.stderr, r.stderr[:4000]            # -fsanitize=alignment: a misaligned float2 / float4 access faults on the device


def test_block_level_kernels_under_address_sanitizer(tmp_path):
    """The same sweep under AddressSanitizer (the launch's dynamic shared memory is a heap block of exactly the requested size, the global
    buffers are numpy's): no access outside either."""
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    rt = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(rt) or not os.path.exists(rt):
        pytest.skip("no AddressSanitizer runtime")
    _build_all(tmp_path, ["-fno-gnu-unique", "-fsanitize=address"])
    env = dict(os.environ, LD_PRELOAD=rt, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1", B200_EMUL_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "cpu_emul", "tsan_sweep.py"), str(tmp_path)], capture_output=True, text=True,
                       env=env, timeout=900)
    if "EMUL_DONE" not in r.stdout and "AddressSanitizer" not in r.stderr:
        pytest.skip("AddressSanitizer could not run here: %s" % r.stderr[-300:])
    assert "AddressSanitizer" not in r.stderr, r.stderr[:4000]
    assert "EMUL_DONE" in r.stdout, r.stderr[-2000:]
