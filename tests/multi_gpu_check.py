"""Multi-GPU parity check, one process per GPU:

    torchrun --nnodes=1 --nproc-per-node N tests/multi_gpu_check.py

Every rank runs the adaptation pipeline on its GPU through the TMROctForest
API; rank r's octants, conn, node numbers, node_range and dependent CSR are
compared bit for bit with rank r of the reference's own MPI path (oracle,
thread-ranks on the host, same rank count).  Exits non-zero on any mismatch."""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import multirank  # noqa: E402
import util  # noqa: E402


def check_cases(lib, rank, size, verbose=True):
    """Run every multi-rank case defined for `size` ranks on the GPUs (one
    process per GPU, NCCL already initialised) and compare each rank's results
    with the same rank of the oracle run at the same rank count.  Returns
    (cases, failures, cases without an oracle) on every rank."""
    from oracle import ref_loader

    ref = ref_loader.load() if (rank == 0 and ref_loader.available()) else None
    failures = unchecked = 0
    cases = [c for c in multirank.CASES if c[7] == size]
    # a heavier case at this rank count: 2x2x2 trees, 4 passes
    cases.append(("grid2_big_r%d" % size, "grid2", 2, 4, 35, 0, 2, size, True))
    cases.append(("butterfly_o3_r%d" % size, "butterfly2", 1, 2, 30, 1, 3, size, True))
    for case in cases:
        name, conn_name, level, passes, pct, corner, order, ranks, repart = case
        conn = util.CONNS[conn_name]()
        body = multirank.adapt_body(conn, level, passes, pct, corner, order, repart,
                                    with_interp=True)
        try:
            mine = body(lib, rank)
        except Exception:  # noqa: BLE001
            traceback.print_exc()
            mine = None
        gathered = [None] * size
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            try:
                assert all(g is not None for g in gathered), "a rank failed"
                total = sum(len(g[0][-1]) for g in gathered)
                if ref is not None:
                    expect = multirank.run_thread_ranks(ref, size, body, True)
                    multirank.compare_rank_results(expect, gathered, name)
                    how = "bit-exact vs reference at %d ranks" % size
                else:
                    unchecked += 1
                    how = "oracle not available: only ran"
                allocts = np.concatenate([g[0][-1] for g in gathered])
                if verbose:
                    print("[multi-gpu] %-28s OK  %8d octants  checksum %016x  (%s)" %
                          (name, total, util.checksum(allocts), how), file=sys.stderr,
                          flush=True)
            except AssertionError as e:
                failures += 1
                print("[multi-gpu] %-28s FAIL %s" % (name, str(e)[:400]), file=sys.stderr,
                      flush=True)
    # findEnclosing for nodes held by OTHER ranks: no element, mpi_owner = owner rank
    conn = util.box_conn()
    q = [multirank.find_enclosing_queries(ref, conn) if ref is not None else None]
    dist.broadcast_object_list(q, src=0)
    ncases = len(cases)
    if q[0] is not None:
        ncases += 1
        body = multirank.find_enclosing_body(conn, q[0])
        try:
            mine = body(lib, rank)
        except Exception:  # noqa: BLE001
            traceback.print_exc()
            mine = None
        gathered = [None] * size
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            try:
                assert all(g is not None for g in gathered), "a rank failed"
                expect = multirank.run_thread_ranks(ref, size, body, True)
                nm = multirank.compare_find_enclosing(expect, gathered, size)
                if verbose:
                    print("[multi-gpu] %-28s OK  %8d misses name the reference's owner rank" %
                          ("find_enclosing_r%d" % size, nm), file=sys.stderr, flush=True)
            except AssertionError as e:
                failures += 1
                print("[multi-gpu] find_enclosing FAIL %s" % str(e)[:400], file=sys.stderr,
                      flush=True)
    flag = torch.tensor([failures, unchecked], device="cuda")
    dist.broadcast(flag, src=0)
    return ncases, int(flag[0].item()), int(flag[1].item())


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ["TMR_B200_DEVICE"] = str(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import tmr_b200
    from tmr_b200 import dist as tdist

    lib = tmr_b200.require_gpu()
    rank, size = tdist.init_from_torch(lib)
    _, failures, _ = check_cases(lib, rank, size)
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
