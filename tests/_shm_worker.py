"""Worker of tests/test_shm_transport_cpu.py: one rank of the shared-memory transport's CONTROL PLANE (no device
needed): usage  python tests/_shm_worker.py <rank> <world> <id file>"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quest_b200 import capi      # noqa: E402


def main():
    rank, world, idfile = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    lib = capi.lib()
    idbuf = C.create_string_buffer(128)
    if rank == 0:
        capi.check(lib.qb_comm_set_transport(1))
        capi.check(lib.qb_comm_get_unique_id(idbuf))
        with open(idfile + ".tmp", "wb") as f:
            f.write(idbuf.raw)
        os.rename(idfile + ".tmp", idfile)
    else:
        for _ in range(600):
            if os.path.exists(idfile):
                break
            time.sleep(0.05)
        idbuf.raw = open(idfile, "rb").read()
    capi.check(lib.qb_comm_init(rank, world, idbuf), "qb_comm_init")
    assert lib.qb_comm_transport() == 1 and lib.qb_comm_rank() == rank and lib.qb_comm_num_ranks() == world

    # all-reduce (more values than one mail slot holds -> chunked), identical bits on every rank
    n = 300000
    v = (np.arange(n, dtype=np.float64) * (rank + 1) * 0.1)
    capi.check(lib.qb_comm_allreduce_sum(v.ctypes.data_as(C.c_void_p), n))
    want = np.zeros(n)
    for r in range(world):
        want += np.arange(n, dtype=np.float64) * (r + 1) * 0.1
    assert np.array_equal(v, want), "allreduce"

    flag = C.c_int(1 if rank != world - 1 else 0)
    capi.check(lib.qb_comm_allreduce_and(C.byref(flag)))
    assert flag.value == 0
    flag = C.c_int(1)
    capi.check(lib.qb_comm_allreduce_and(C.byref(flag)))
    assert flag.value == 1

    # broadcast from every root, larger than a slot
    for root in range(world):
        b = np.full(200000, float(rank), dtype=np.float64)
        capi.check(lib.qb_comm_broadcast_bytes(b.ctypes.data_as(C.c_void_p), b.nbytes, root))
        assert np.all(b == root), "broadcast"

    # gather to root
    mine = np.frombuffer(("rank%03d" % rank).encode(), dtype=np.uint8).copy()
    allb = np.zeros(mine.size * world, dtype=np.uint8)
    capi.check(lib.qb_comm_gather_bytes(mine.ctypes.data_as(C.c_void_p), allb.ctypes.data_as(C.c_void_p), mine.size, 0))
    if rank == 0:
        assert allb.tobytes() == b"".join(("rank%03d" % r).encode() for r in range(world)), "gather"

    # host amplitudes from every rank to the root (comm_sendAmpsToRoot), only the two ranks involved take part
    for sender in range(1, world):
        send = (np.arange(70000) + 1j * sender).astype(np.complex128)
        recv = np.zeros_like(send)
        capi.check(lib.qb_comm_sendrecv_host(send.ctypes.data_as(C.c_void_p), recv.ctypes.data_as(C.c_void_p), send.size, sender, 0))
        if rank == 0:
            assert np.array_equal(recv, send), "sendrecv_host"

    # device data plane must fail loudly without a device (no CPU fallback)
    st = lib.qb_comm_exchange(None, None, 16, rank ^ 1)
    assert st != 0 and b"no CPU fallback" in lib.qb_error_string(), lib.qb_error_string()

    capi.check(lib.qb_comm_barrier())
    capi.check(lib.qb_comm_end())
    print("ok", rank)


if __name__ == "__main__":
    main()
