"""The "ranks share a GPU" transport (quest_b200/csrc/qb_comm_shm.cu): its host control plane -- rendezvous,
all-reduce, broadcast, gather, point-to-point host messages -- runs here with 2 and 4 processes and no device.
Its device data plane (CUDA IPC staging) is what tests/test_dist_gpu.py runs on when the box has fewer GPUs
than ranks."""
import os
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.parametrize("world", [2, 4])
def test_shared_memory_control_plane(world):
    if _have_gpu():
        pytest.skip("control-plane-only test (expects the data plane to refuse for lack of a device)")
    with tempfile.TemporaryDirectory() as d:
        idfile = os.path.join(d, "id")
        procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_shm_worker.py"), str(r), str(world), idfile],
                                  stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(world)]
        for r, p in enumerate(procs):
            try:
                so, se = p.communicate(timeout=120)
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                raise
            assert p.returncode == 0 and f"ok {r}" in so, f"rank {r}: rc={p.returncode}\n{so}\n{se[-2000:]}"


def test_dead_partner_is_an_error_not_a_hang():
    """a rank that waits for a partner that never joins gives up after QUEST_B200_SHM_TIMEOUT_S with an error"""
    code = (
        "import ctypes as C, sys; sys.path.insert(0, %r)\n"
        "from quest_b200 import capi\n"
        "lib = capi.lib(); b = C.create_string_buffer(128)\n"
        "lib.qb_comm_set_transport(1); capi.check(lib.qb_comm_get_unique_id(b))\n"
        "st = lib.qb_comm_init(0, 2, b)\n"
        "print('status', st, lib.qb_error_string().decode())\n" % ROOT)
    env = dict(os.environ, QUEST_B200_SHM_TIMEOUT_S="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "status 0" not in r.stdout and "timed out" in r.stdout, r.stdout
