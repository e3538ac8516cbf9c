"""world_size-2 gloo test (CPU) of the single collective of the path: the fixed-shape detection-record all-gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from sipmask_b200 import dist as D
    g = torch.Generator().manual_seed(rank)
    max_num, k = 10, 3 + 4 * rank
    det = torch.zeros(max_num, 5)
    det[:k] = torch.rand(k, 5, generator=g) + rank
    lab = torch.full((max_num,), -1, dtype=torch.long)
    lab[:k] = torch.arange(k) + 10 * rank
    rec = D.pack_record(det, lab, torch.tensor([k], dtype=torch.int32))
    out = D.gather_records(rec)
    res = D.unpack_records(out)
    ok = len(res) == world
    for r, (b, l) in enumerate(res):
        kk = 3 + 4 * r
        gg = torch.Generator().manual_seed(r)
        exp = torch.rand(kk, 5, generator=gg) + r
        ok = ok and b.shape == (kk, 5) and torch.allclose(b, exp) and l.tolist() == (torch.arange(kk) + 10 * r).tolist()
    # the run-level log: K images appended locally, ONE collective at the end
    log = D.RecordLog(4, max_num, 'cpu')
    for i in range(3):
        log.append(det + i, lab, torch.tensor([k], dtype=torch.int32))
    allrec = log.gather()
    ok = ok and tuple(allrec.shape) == (world, 4, max_num, 7)
    for r in range(world):
        kk = 3 + 4 * r
        gg = torch.Generator().manual_seed(r)
        exp = torch.rand(kk, 5, generator=gg) + r
        for i in range(3):
            ok = ok and torch.allclose(allrec[r, i, :kk, :5], exp + i) and int(allrec[r, i, :, 6].sum()) == kk
    # instance masks: RLE run-length tensors gathered as they are
    counts = torch.full((max_num, 6), rank, dtype=torch.int32)
    ncnt = torch.full((max_num,), 2 + rank, dtype=torch.int32)
    gc, gn = D.gather_rle(counts, ncnt)
    ok = ok and gc.shape == (world, max_num, 6) and all(int(gc[r].max()) == r and int(gn[r][0]) == 2 + r for r in range(world))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_detection_record_all_gather_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]


def test_single_process_passthrough():
    from sipmask_b200 import dist as D
    det = torch.rand(5, 5)
    lab = torch.arange(5)
    rec = D.pack_record(det, lab, torch.tensor([4], dtype=torch.int32))
    out = D.unpack_records(D.gather_records(rec))
    assert len(out) == 1 and out[0][0].shape == (4, 5) and out[0][1].tolist() == [0, 1, 2, 3]
