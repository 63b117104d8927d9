"""N > 1 host logic on CPU: world_size-2 gloo run of the utterance sharding + the single waveform gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_utt, q):
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from mlx_audio_swift_b200 import dist as bd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bd.shard_utterances(n_utt, rank, world)
    t_max = 16
    local = torch.stack([torch.full((t_max,), float(u)) + torch.arange(t_max) * 0.01 for u in mine]) if mine else torch.zeros((0, t_max))
    lens = torch.as_tensor([4 + u for u in mine], dtype=torch.int64)
    waves, wl = bd.gather_waveforms(local, lens, n_utt)
    ok = waves.shape == (n_utt, t_max) and all(abs(float(waves[u, 0]) - u) < 1e-6 for u in range(n_utt)) \
        and wl.tolist() == [4 + u for u in range(n_utt)]
    q.put((rank, ok, mine))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_utt", [8, 5])
def test_shard_and_gather_world2(n_utt):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_utt, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    shards = {r: m for r, _, m in res}
    assert sorted(shards[0] + shards[1]) == list(range(n_utt)) and not set(shards[0]) & set(shards[1])


def test_shard_is_a_partition():
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    from mlx_audio_swift_b200.dist import shard_utterances
    for n in (0, 1, 7, 8, 64):
        for w in (1, 2, 4, 8):
            parts = [shard_utterances(n, r, w) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard_utterances(4, 2, 2)
