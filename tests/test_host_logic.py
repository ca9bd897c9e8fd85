"""CPU: host-side logic — golden fixtures vs the compiled reference, synthetic generator
determinism, utterance sharding and the world_size-2 statistics reduce over gloo."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN_NAMES, ROOT, load_golden
from oracle import metrics as M


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_golden_vectors_are_the_reference(reference_lib, name):
    """Pins the fixtures: re-running the unmodified reference reproduces them."""
    g = load_golden(name)
    x = g["pcm"].astype(np.float64) / 32768.0
    fs = int(g["fs"])
    o = reference_lib.analyze(x, fs)
    assert o["fft_size"] == int(g["fft_size"])
    assert np.array_equal(o["t"], g["t"])
    assert np.array_equal(o["f0_raw"], g["f0_raw"])
    assert np.array_equal(o["f0"], g["f0"])
    assert np.array_equal(o["sp"][g["rows"]].astype(np.float32), g["sp_rows"])
    assert np.array_equal(o["ap"][g["rows"]].astype(np.float32), g["ap_rows"])
    y = reference_lib.synthesis(o["f0"], o["sp"], o["ap"], o["fft_size"], 5.0, fs)
    assert np.array_equal(y.astype(np.float32), g["y"])


def test_randn_golden_is_the_reference(reference_lib):
    from oracle import ref
    gold = np.load(os.path.join(ROOT, "tests", "golden", "randn_first_8192.npy"))
    assert np.array_equal(ref.randn_stream(8192), gold)


def test_metrics_definitions():
    a = np.array([0.0, 100.0, 200.0, 0.0])
    b = np.array([0.0, 100.01, 0.0, 0.0])
    assert M.vuv_agreement(a, b) == 0.75
    assert np.isclose(M.f0_rel_error(a, b), 1e-4)
    sp = np.ones((3, 5))
    assert M.lsd_db(sp, sp * 10 ** 0.1)[1] == pytest.approx(1.0)
    assert M.ap_abs_error(sp, sp + 0.5) == 0.5
    y = np.sin(np.arange(1000.0))
    assert M.snr_db(y, y * (1 + 1e-3)) == pytest.approx(60.0, abs=1e-6)


def test_signal_generator_is_deterministic_and_speechlike():
    from hts_train_world_b200 import signals
    a, pa = signals.make_utterance(42, 16000)
    b, pb = signals.make_utterance(42, 16000)
    assert np.array_equal(a.numpy(), b.numpy()) and pa == pb
    assert 1.5 * 16000 <= len(a) <= 6.0 * 16000
    assert a.dtype.is_floating_point is False and int(a.abs().max()) > 15000
    assert 80.0 <= pa["f0_base"] <= 300.0
    c, _ = signals.make_utterance(43, 16000)
    assert len(c) != len(a) or not np.array_equal(a.numpy(), c.numpy())
    assert (a.numpy() != 0).mean() > 0.95          # noise floor everywhere: no digital silence


def test_sharding_is_a_balanced_partition():
    from hts_train_world_b200 import corpus
    rng = np.random.default_rng(0)
    lengths = rng.integers(72000, 288000, size=1132)
    for world in [1, 2, 4, 8]:
        parts = [corpus.shard_utterances(lengths, r, world) for r in range(world)]
        allu = np.sort(np.concatenate(parts))
        assert np.array_equal(allu, np.arange(len(lengths)))
        loads = np.array([lengths[p].sum() for p in parts])
        assert loads.max() - loads.min() <= lengths.max()


def test_merge_stats():
    from hts_train_world_b200 import corpus
    rng = np.random.default_rng(1)
    v = rng.normal(5.0, 0.3, size=1000)
    parts = [[len(c), c.sum(), (c * c).sum()] for c in np.array_split(v, 7)]
    m = corpus.merge_stats(parts)
    assert m["count"] == 1000 and np.isclose(m["mean"], v.mean()) and np.isclose(m["var"], v.var())


_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, %r)
import torch.distributed as dist
import hts_train_world_b200
from hts_train_world_b200 import corpus
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% sys.argv[1],
                        rank=int(sys.argv[2]), world_size=2)
rank = dist.get_rank()
rng = np.random.default_rng(7)
lengths = rng.integers(1000, 5000, size=40)
mine = corpus.shard_utterances(lengths, rank, 2)
vals = [np.log(100.0 + u + np.arange(lengths[u] // 100)) for u in mine]
loc = np.array([sum(len(v) for v in vals), sum(v.sum() for v in vals), sum((v * v).sum() for v in vals)])
loc_mine = loc.copy()                       # (allreduce_stats reduces in place)
m = corpus.allreduce_stats(loc)
allv = np.concatenate([np.log(100.0 + u + np.arange(lengths[u] // 100)) for u in range(40)])
assert m["count"] == len(allv), (m, len(allv))
assert np.isclose(m["mean"], allv.mean()) and np.isclose(m["var"], allv.var())
# the corpus driver's single all-reduce of all partials (feature statistics + global-variance partials):
# every rank holds the per-utterance variances of its shard, the merged result is the variance of all of them
from hts_train_world_b200 import driver
per = np.stack([np.array([np.var(np.sin(np.arange(50 + u) * (0.1 + 0.01 * k))) for k in range(4)]) for u in range(40)])
mine_v = per[mine].astype(np.float32).astype(np.float64)
gv_part = np.stack([np.full(4, float(len(mine_v))), mine_v.sum(axis=0), (mine_v * mine_v).sum(axis=0)], axis=1)
stats_part = np.stack([loc_mine, loc_mine * 2.0])
got = driver.all_reduce_partials([stats_part, gv_part])
assert got[0].shape == (2, 3) and np.isclose(got[0][0, 0], len(allv)) and np.isclose(got[0][1, 1], 2.0 * allv.sum())
mean, var = corpus.merge_gv([got[1]])
allp = per.astype(np.float32).astype(np.float64)
assert np.allclose(mean, allp.mean(axis=0)) and np.allclose(var, allp.var(axis=0), rtol=1e-9, atol=1e-18)
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_stats_allreduce_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = [subprocess.Popen([sys.executable, str(script), str(port), str(r)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % r in o


def test_driver_host_rules():
    """clip check of data/Makefile.in:127-129 and FFTLEN of configure.ac:540-549."""
    from hts_train_world_b200 import driver
    ok = np.array([1, -5, 32766, -32767], np.int16)
    assert driver.passes_clip_check(ok)
    assert not driver.passes_clip_check(np.array([0, 32767], np.int16))
    assert not driver.passes_clip_check(np.array([-32768, 3], np.int16))
    assert not driver.passes_clip_check(np.zeros(0, np.int16))
    assert [driver.fftlen_for(fs) for fs in (16000, 22050, 44100, 48000)] == [1024, 1024, 2048, 2048]


def test_reference_arm_deals_utterances_evenly():
    """bench.py's CPU arm: every core gets about the same seconds of audio (longest first onto the least loaded core),
    every utterance exactly once."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from hts_train_world_b200 import signals
    utts = bench.reference_sample_utts(16, 4)
    bins = bench.deal_utterances(utts, 16)
    assert sorted(u for b in bins for u in b) == utts
    loads = [sum(signals.utterance_params(u)["T"] for u in b) for b in bins]
    assert max(loads) <= 1.15 * (sum(loads) / len(loads))
