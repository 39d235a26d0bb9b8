"""Input pipeline: oracle/input_pipeline.py pinned to torchvision / Pillow and to the golden vectors
the reference's own MonoDataset.preprocess / KITTIInpaintDataset.preprocess_masks produced; the CUDA path (tdl_input_fwd)
bit-exact against the oracle."""
import itertools
import os

import numpy as np
import pytest
import torch

from oracle import input_pipeline as oip

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "input", "input_b3_24x40.pt")


def _all_colors(stride=1):
    v = np.arange(0, 1 << 24, stride, dtype=np.uint32)
    return np.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], -1).astype(np.uint8).reshape(1, -1, 3)


def _oracle_from_golden(rec):
    fr = [rec["frames_u8"][:, i].numpy() for i in range(len(rec["frame_ids"]))]
    return oip.input_pipeline(fr, rec["jitter"].numpy(), rec["order"].numpy(), rec["do_aug"].numpy().astype(bool), None,
                              rec["holes"].numpy(), *rec["erase_shape"])


def test_oracle_reproduces_the_reference_golden_vectors_bit_exactly():
    rec = torch.load(GOLD)
    color, aug, mask = _oracle_from_golden(rec)
    for i in range(len(rec["frame_ids"])):
        assert np.array_equal(color[i], rec["color"][i].numpy())
        assert np.array_equal(aug[i], rec["color_aug"][i].numpy())
    assert np.array_equal(mask, rec["mask"].numpy())
    assert float((rec["color"] != rec["color_aug"]).float().mean()) > 0.3       # the augmentation did something
    for b in range(rec["jitter"].shape[0]):                                   # the byte the hue step adds
        for f in range(rec["jitter"].shape[1]):
            if rec["do_aug"][b]:
                assert oip.hue_shift_u8(float(rec["hue_factor"][b, f])) == int(rec["jitter"][b, f, 3])


def test_oracle_matches_pillow_and_torchvision_per_operation():
    PIL = pytest.importorskip("PIL")
    from PIL import Image
    import torchvision.transforms.functional as TF
    rng = np.random.default_rng(5)
    # every 5th colour of the RGB cube (3.4 M pixels) + a textured image for the image-level mean of the contrast step
    cube = _all_colors(5)
    tex = (rng.random((64, 96, 3)) ** 2 * 255).astype(np.uint8)
    for arr in (cube, tex):
        img = Image.fromarray(arr)
        assert np.array_equal(np.array(img.convert("L")), oip.gray_u8(arr))
        hsv = np.array(img.convert("HSV"))
        assert np.array_equal(hsv, oip.rgb_to_hsv_u8(arr))
        assert np.array_equal(np.array(Image.fromarray(hsv, "HSV").convert("RGB")), oip.hsv_to_rgb_u8(hsv))
        for f in (0.8, 0.8731, 1.0, 1.1234, 1.2, 0.0, 2.5):
            assert np.array_equal(np.array(TF.adjust_brightness(img, f)), oip.adjust_brightness(arr, f)), f
            assert np.array_equal(np.array(TF.adjust_contrast(img, f)), oip.adjust_contrast(arr, f)), f
            assert np.array_equal(np.array(TF.adjust_saturation(img, f)), oip.adjust_saturation(arr, f)), f
        for h in (-0.1, -0.0371, 0.0, 0.0039, 0.1, 0.5, -0.5):
            assert np.array_equal(np.array(TF.adjust_hue(img, h)), oip.adjust_hue(arr, h)), h
    # HSV -> RGB over every 3rd HSV triple (also triples that RGB -> HSV never produces)
    hsv = _all_colors(3)
    assert np.array_equal(np.array(Image.fromarray(hsv, "HSV").convert("RGB")), oip.hsv_to_rgb_u8(hsv))
    assert np.array_equal(oip.to_tensor(tex), TF.to_tensor(Image.fromarray(tex)).numpy())


def test_oracle_color_jitter_is_torchvisions_on_pil_images():
    pytest.importorskip("PIL")
    from PIL import Image
    from torchvision import transforms
    rng = np.random.default_rng(6)
    arr = (rng.random((48, 80, 3)) * 255).astype(np.uint8)
    img = Image.fromarray(arr)
    cj = transforms.ColorJitter((0.8, 1.2), (0.8, 1.2), (0.8, 1.2), (-0.1, 0.1))
    for seed in range(12):
        torch.manual_seed(seed)
        fn_idx, b, c, s, h = cj.get_params(cj.brightness, cj.contrast, cj.saturation, cj.hue)
        torch.manual_seed(seed)
        ref = np.array(cj(img))
        mine = oip.color_jitter(arr, fn_idx.tolist(), b, c, s, oip.hue_shift_u8(h))
        assert np.array_equal(ref, mine), (seed, fn_idx)


def test_sample_params_ranges(tdl):
    from importlib import import_module
    ip = import_module(tdl.__name__ + ".input_pipeline")
    pipe = ip.GpuInputPipeline([0, -1, 1], 192, 640, erase_count=16, erase_shape=(16, 16))
    g = torch.Generator().manual_seed(1)
    P = pipe.sample_params(64, g)
    j = P["jitter"]
    assert j.shape == (64, 3, 4) and float(j[..., :3].min()) >= 0.8 and float(j[..., :3].max()) <= 1.2
    hb = j[..., 3]
    assert bool(((hb <= 25) | (hb >= 231)).all())                    # uint8(int32(h * 255)), |h| <= 0.1
    assert sorted(P["order"][5, 1].tolist()) == [0, 1, 2, 3]
    assert 0.2 < float(P["do_aug"].float().mean()) < 0.8 and 0.2 < float(P["do_flip"].float().mean()) < 0.8
    assert int(P["holes"][..., 0].max()) < 192 - 16 - 1 and int(P["holes"][..., 1].max()) < 640 - 16 - 1
    with pytest.raises(tdl._lib.TdlError):                           # no CPU path
        pipe({f: torch.zeros(2, 192, 640, 3, dtype=torch.uint8) for f in (0, -1, 1)})


# ------------------------------------------------------------------------------------------------ GPU
def _run_gpu(tdl, frames, params, frame_ids, erase, want_aug=True):
    from importlib import import_module
    ip = import_module(tdl.__name__ + ".input_pipeline")
    B, H, W, _ = frames[0].shape
    pipe = ip.GpuInputPipeline(frame_ids, H, W, erase_count=(params["holes"].shape[1] if params.get("holes") is not None else 0),
                               erase_shape=erase)
    out = pipe({f: torch.from_numpy(np.ascontiguousarray(frames[i])).cuda() for i, f in enumerate(frame_ids)},
               {k: (torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v) for k, v in params.items()},
               want_color_aug=want_aug)
    torch.cuda.synchronize()
    return out


@pytest.mark.gpu
def test_gpu_matches_the_reference_golden_vectors_bit_exactly(tdl):
    rec = torch.load(GOLD)
    fids = rec["frame_ids"]
    frames = [rec["frames_u8"][:, i].numpy() for i in range(len(fids))]
    params = dict(jitter=rec["jitter"], order=rec["order"], do_aug=rec["do_aug"], do_flip=None, holes=rec["holes"])
    out = _run_gpu(tdl, frames, params, fids, rec["erase_shape"])
    for i, f in enumerate(fids):
        assert torch.equal(out[("color", f, 0)].cpu(), rec["color"][i])
        assert torch.equal(out[("color_aug", f, 0)].cpu(), rec["color_aug"][i])
    assert torch.equal(out[("mask", 0, 0)].cpu(), rec["mask"])


@pytest.mark.gpu
def test_gpu_matches_the_oracle_on_a_training_batch(tdl):
    rng = np.random.default_rng(11)
    B, H, W, fids = 4, 192, 640, [0, -1, 1]
    tex = rng.random((B, 3, H, W, 3))
    frames = [((tex[:, i] ** (1 + i)) * 255).astype(np.uint8) for i in range(3)]
    perms = list(itertools.permutations(range(4)))
    order = np.array([[perms[(7 * b + 5 * f + 3) % 24] for f in range(3)] for b in range(B)], np.int32)
    jitter = rng.uniform(0.8, 1.2, (B, 3, 4)).astype(np.float32)
    jitter[..., 3] = rng.integers(0, 256, (B, 3))
    jitter[0, 0, :3] = (1.7, 0.0, 2.2)                               # outside [0, 1]: the clipping branch of Image.blend
    do_aug = np.array([1, 1, 0, 1], np.uint8)
    do_flip = np.array([0, 1, 1, 0], np.uint8)
    holes = np.stack([rng.integers(0, H - 17, (B, 16)), rng.integers(0, W - 17, (B, 16))], -1).astype(np.int32)
    params = dict(jitter=jitter, order=order, do_aug=do_aug, do_flip=do_flip, holes=holes)
    out = _run_gpu(tdl, frames, params, fids, (16, 16))
    color, aug, mask = oip.input_pipeline(frames, jitter, order, do_aug.astype(bool), do_flip.astype(bool), holes, 16, 16)
    for i, f in enumerate(fids):
        assert np.array_equal(out[("color", f, 0)].cpu().numpy(), color[i])
        assert np.array_equal(out[("color_aug", f, 0)].cpu().numpy(), aug[i])
    assert np.array_equal(out[("mask", 0, 0)].cpu().numpy(), mask)
    assert 0.03 < 1 - float(mask.mean()) < 0.04                      # 16 boxes of 16x16 on 192x640, overlaps aside
    # no jitter requested: color only, no workspace
    out2 = _run_gpu(tdl, frames, dict(jitter=None, order=None, do_aug=None, do_flip=do_flip, holes=None), fids, (16, 16), want_aug=False)
    assert np.array_equal(out2[("color", 0, 0)].cpu().numpy(), color[0]) and ("color_aug", 0, 0) not in out2


@pytest.mark.gpu
def test_gpu_hue_and_blends_over_the_whole_colour_cube(tdl):
    """Every RGB triple through the device arithmetic (float / double mix of Pillow's Convert.c, Blend.c)."""
    cube = _all_colors().reshape(1, 4096, 4096, 3)
    for order, jit in (((3, 9, 9, 9), (1.0, 1.0, 1.0, 231.0)), ((2, 0, 9, 9), (1.19, 1.0, 0.83, 0.0)),
                       ((1, 3, 2, 0), (0.91, 1.13, 1.07, 17.0))):
        jitter = np.array(jit, np.float32).reshape(1, 1, 4)
        o = np.array(order, np.int32).reshape(1, 1, 4)
        out = _run_gpu(tdl, [cube], dict(jitter=jitter, order=o, do_aug=np.ones(1, np.uint8), do_flip=None, holes=None), [0], (1, 1))
        ref = oip.to_tensor(oip.color_jitter(cube[0], [k for k in order if k < 4], jit[0], jit[1], jit[2], int(jit[3])))
        got = out[("color_aug", 0, 0)][0].cpu().numpy()
        assert np.array_equal(got, ref), (order, jit, int((got != ref).sum()))
