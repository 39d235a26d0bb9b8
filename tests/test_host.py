"""CPU: host-side logic -- config shim, net table, pose helpers, synthetic inputs, bench helpers and the
2-rank (gloo) path of bench.py's sharding / max-over-ranks reduction."""
import importlib
import os
import sys
import textwrap

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_config_fromfile(tdl, tmp_path):
    cfg_py = tmp_path / "cfg.py"
    cfg_py.write_text(textwrap.dedent('''
        FRAME_IDS = [0, -1, 1]
        model = dict(name='mono_fm', frame_ids=FRAME_IDS, height=192, width=640, scales=[0, 1, 2, 3],
                     automask=False if 's' in FRAME_IDS else True, perception_weight=1e-3)
        optimizer = dict(type='Adam', lr=1e-4)
        dist_params = dict(backend='nccl')
    '''))
    cfg = tdl.Config.fromfile(str(cfg_py))
    assert cfg.model.name == "mono_fm" and cfg.model["height"] == 192
    assert cfg.model.get("missing", 7) == 7 and cfg.model.automask is True
    assert cfg.dist_params.backend == "nccl"
    with pytest.raises(AttributeError):
        cfg.model.nope


def test_net_table(tdl):
    import torch.nn as nn
    table = tdl.registry.NetTable("t")

    @table.register_module
    class A(nn.Module):
        def __init__(self, opt):
            super().__init__()
            self.opt = opt

    assert table.module_dict["A"] is A and table.name == "t"
    assert isinstance(table.build(dict(name="A")), A)
    with pytest.raises(KeyError):
        table.register_module(A)
    with pytest.raises(TypeError):
        table.register_module(int)


def test_pose_helpers_match_oracle(tdl):
    from oracle import restatement as R
    g = torch.Generator().manual_seed(0)
    aa, tr = 0.1 * torch.randn(3, 1, 3, generator=g), torch.randn(3, 1, 3, generator=g)
    for inv in (False, True):
        torch.testing.assert_close(tdl.geometry.transformation_from_parameters(aa, tr, inv),
                                   R.transformation_from_parameters(aa, tr, inv))
    K, iK = tdl.synth.kitti_intrinsics(2, 192, 640)
    Kh, inv = tdl.geometry.half_res_intrinsics(K, iK)
    Kh_ref, inv_ref = R.half_res_intrinsics(K)
    torch.testing.assert_close(Kh, Kh_ref, rtol=0, atol=0)
    torch.testing.assert_close(inv, inv_ref, rtol=1e-6, atol=1e-6)


def test_synth_is_deterministic(tdl):
    a = tdl.synth.make_inputs(1, 32, 64, seed=5, feat_channels=4)
    b = tdl.synth.make_inputs(1, 32, 64, seed=5, feat_channels=4)
    for da, db in zip(a[:2], b[:2]):
        for k in da:
            assert torch.equal(da[k], db[k])
    c = tdl.synth.make_inputs(1, 32, 64, seed=6)
    assert not torch.equal(a[0][("color", 0, 0)], c[0][("color", 0, 0)])
    img = a[0][("color", 0, 0)]
    assert img.min() >= 0 and img.max() <= 1 and a[1][("disp", 0, 3)].shape == (1, 1, 2, 4)


def test_loss_dict_total(tdl):
    """total() == sum of .mean() of every entry (mono/apis/trainer.py:39-48), whether an entry is covered by a
    fused-kernel part or was appended the reference's way."""
    d = tdl.losses.LossDict()
    v = torch.tensor([1.0, 2.0, 3.0])
    d["a"], d["b"], d["c"] = v[0], v[1], v[2]
    d.add_part(v, 1, ["a", "b", "c"])
    per = torch.tensor(0.5)
    for s in range(4):
        d[("p", s)] = per
    d.add_part(per, 4, [("p", s) for s in range(4)])
    assert float(d.total()) == 8.0
    # entries added like the reference does (loss_dict.update(self.compute_xxx_loss(...)), loss_dict[k] = v)
    d.update({"extra": torch.tensor([[1.0, 3.0]])})          # un-reduced map: batch_processor takes its mean
    d["more"] = torch.tensor(0.25)
    assert float(d.total()) == 8.0 + 2.0 + 0.25
    assert float(d.total()) == float(sum(x.mean() for x in d.values()))
    # overwriting a covered entry drops the part; the remaining views are summed one by one
    d["a"] = torch.tensor(10.0)
    assert float(d.total()) == float(sum(x.mean() for x in d.values()))
    del d["b"]
    assert float(d.total()) == float(sum(x.mean() for x in d.values()))
    e = tdl.losses.LossDict()
    e["x"] = torch.tensor(2.0)
    assert float(e.total()) == 2.0
    f = tdl.losses.LossDict()
    f["y"] = torch.tensor(1.0)
    f.absorb(d)
    assert float(f.total()) == 1.0 + float(d.total())


def test_algorithmic_bytes_match_survey():
    import bench
    b = bench.algorithmic_bytes(1, 192, 640, 2, 64, 4, True, materialize=True, noise_tensors=True)
    N = 192 * 640
    assert b["photo_fwd"] == 304 * N + 4 * sum(N // 4 ** (s + 1) for s in range(4))
    assert b["photo_bwd"] == 176 * N + 8 * sum(N // 4 ** (s + 1) for s in range(4))
    assert b["feat_fwd"] == 320 * N
    assert b["feat_bwd"] + b["memset_dsrc"] == 512 * N
    assert bench.algorithmic_bytes(1, 192, 640, 2, 64, 4, False)["feat_bwd"] == 192 * N


def _rank_main(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ms = bench.max_over_ranks(10.0 + 5.0 * rank, torch.device("cpu"), True)
    seed = bench.rank_seed(rank)
    dist.barrier()
    dist.destroy_process_group()
    out.put((rank, ms, seed))


def test_two_rank_sharding_and_max_reduction():
    """world_size 2 over gloo: ranks draw different shards, the timed figure is the max over ranks and the
    reported value counts the images of ALL ranks."""
    import bench
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [15.0, 15.0]
    assert res[0][2] != res[1][2]
    assert bench.whole_job_images_per_s(2, 8, 10, 15.0) == pytest.approx(2 * 8 * 10 / 0.015)
