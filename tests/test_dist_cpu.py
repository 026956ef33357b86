"""world_size-2 gloo tests (CPU) of the residue-sharding plumbing in
cuhe_b200/sharded.py: the all-gather before ICRT reassembles prime order, the
coefficient-sliced RAW all-gather reassembles polynomials, and the modSwitch
broadcast delivers the dropped residue -- checked against oracle data."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import ROOT, SIMPLE_DHS, get_oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import random
        from cuhe_b200 import sharded as sh
        o = get_oracle(SIMPLE_DHS)
        lvl = 0
        L, H, W = o.L(lvl), o.H, o.W(lvl)
        rng = random.Random(3)
        B = 2
        polys = [[rng.randrange(o.moduli[lvl]) for _ in range(o.n)] for _ in range(B)]
        raws = np.stack([o.to_raw(p, lvl) for p in polys])
        full = np.stack([o.crt(r, lvl) for r in raws])                  # [B][L][H]
        rows_pad = (L + world - 1) // world
        mine = sh.local_primes(L, rank, world)
        assert len(mine) == sh.rows_of(L, rank, world)
        local = np.zeros((B, rows_pad, H), dtype=np.uint32)
        local[:, :len(mine)] = full[:, mine]
        t = torch.from_numpy(local.view(np.int32))
        got = sh.all_gather_residues(t, L, world).numpy().view(np.uint32)
        assert np.array_equal(got, full), "all-gather does not restore prime order"
        # ICRT split by coefficient range (oracle as the compute), RAW slices gathered
        b, e = sh.coefficient_slice(H, rank, world)
        raw_part = np.zeros((B, H, W), dtype=np.uint32)
        for i in range(B):
            raw_part[i, b:e] = o.icrt(got[i], lvl)[b:e]
        raw_all = sh.all_gather_raw(torch.from_numpy(raw_part.view(np.int32)), rank, world).numpy().view(np.uint32)
        assert np.array_equal(raw_all, raws), "RAW all-gather does not restore the polynomials"
        # all-to-all form: each rank receives only its coefficient slice of every residue
        Hs = H // world
        sl = sh.exchange_for_icrt(t, L, rank, world).numpy().view(np.uint32)
        assert sl.shape == (B, L, Hs) and np.array_equal(sl, full[:, :, rank * Hs:(rank + 1) * Hs])
        raw_sl = np.zeros((B, Hs, W), dtype=np.uint32)
        for i in range(B):
            raw_sl[i] = o.icrt(got[i], lvl)[rank * Hs:(rank + 1) * Hs]
        raw_all2 = sh.all_gather_raw_slices(torch.from_numpy(raw_sl.view(np.int32)), world).numpy().view(np.uint32)
        assert np.array_equal(raw_all2, raws)
        # data-parallel front and back: rank r uploads products [r*b, (r+1)*b), everybody gets all of
        # them; after the sliced ICRT every rank gets back the complete results of ITS products only
        Bw = 2 * world
        allp = np.arange(Bw * H * W, dtype=np.uint32).reshape(Bw, H, W) * np.uint32(2654435761)
        own = allp[rank * 2:(rank + 1) * 2]
        gathered = sh.all_gather_operands(torch.from_numpy(own.view(np.int32).copy()), world).numpy().view(np.uint32)
        assert np.array_equal(gathered, allp)
        my_slice = np.ascontiguousarray(allp[:, rank * Hs:(rank + 1) * Hs])          # slice `rank` of every product
        back = sh.raw_slices_to_owners(torch.from_numpy(my_slice.view(np.int32).copy()), world).numpy().view(np.uint32)
        assert np.array_equal(back, own), "slice-to-owner exchange does not rebuild the owned products"
        # modswitch: owner of the last prime broadcasts its row
        owner, row = sh.owner_of(L - 1, world)
        assert owner == (L - 1) % world and sh.local_primes(L, owner, world)[row] == L - 1
        last = sh.broadcast_last_row(torch.from_numpy(local[0].view(np.int32)), L, rank, world).numpy().view(np.uint32)
        assert np.array_equal(last, full[0, L - 1])
        q.put((rank, "ok"))
    except Exception as ex:  # pragma: no cover
        q.put((rank, repr(ex)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharding_plumbing_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_partition_properties():
    from cuhe_b200 import sharded as sh
    for L in (1, 7, 24, 25, 64):
        for G in (1, 2, 4, 8):
            owned = [sh.local_primes(L, r, G) for r in range(G)]
            assert sorted(sum(owned, [])) == list(range(L))
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
            for p in range(L):
                r, i = sh.owner_of(p, G)
                assert owned[r][i] == p
            # dropping the last prime keeps every shard a prefix of itself
            for r in range(G):
                assert sh.local_primes(L - 1, r, G) == [p for p in owned[r] if p < L - 1]
