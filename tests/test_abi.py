"""CPU: the C-ABI library loads, exports every symbol include/tdl.h declares, the ctypes mirrors of the
argument structs have the C layout, and argument errors are reported before anything is launched."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tdl.h")


@pytest.fixture(scope="module")
def lib(tdl):
    import __graft_entry__
    __graft_entry__.build()
    return tdl._lib.lib()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tdl_[a-z_0-9]+)\s*\(", src)))


def test_exports_every_declared_symbol(lib, tdl):
    names = declared_functions()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tdl.h but not exported by libtdl.so"
    assert sorted(tdl._lib.EXPORTS) == names
    assert lib.tdl_abi_version() == tdl._lib.TDL_ABI_VERSION == 4
    assert lib.tdl_strerror(0) == b"ok"
    assert b"NULL" in lib.tdl_strerror(-1)


def test_ctypes_structs_match_c_layout(tdl, tmp_path):
    prog = tmp_path / "layout.c"
    prog.write_text('''#include <stdio.h>
#include <stddef.h>
#include "tdl.h"
int main(void) {
  printf("%zu %zu %zu %zu\\n", sizeof(tdl_photo_args), sizeof(tdl_feat_args), sizeof(tdl_edge_args), sizeof(tdl_kernel_time));
  printf("%zu %zu %zu %zu %zu\\n", offsetof(tdl_photo_args, min_depth), offsetof(tdl_photo_args, target),
         offsetof(tdl_photo_args, noise), offsetof(tdl_photo_args, losses), offsetof(tdl_photo_args, dP));
  printf("%zu %zu %zu %zu %zu\\n", offsetof(tdl_feat_args, tgt), offsetof(tdl_feat_args, loss), offsetof(tdl_feat_args, dP),
         offsetof(tdl_feat_args, layout), offsetof(tdl_feat_args, dtype));
  printf("%zu %zu\\n", offsetof(tdl_edge_args, feature), offsetof(tdl_edge_args, d_feature));
  printf("%zu %zu %zu\\n", sizeof(tdl_recon_args), offsetof(tdl_recon_args, pred), offsetof(tdl_recon_args, d_pred));
  printf("%zu %zu %zu %zu\\n", sizeof(tdl_pose_args), offsetof(tdl_pose_args, axisangle), offsetof(tdl_pose_args, dT),
         offsetof(tdl_feat_args, bwd_scratch));
  printf("%zu %zu %zu %zu %zu\\n", sizeof(tdl_input_args), offsetof(tdl_input_args, frames), offsetof(tdl_input_args, jitter),
         offsetof(tdl_input_args, color_aug), offsetof(tdl_input_args, workspace_bytes));
  return 0; }''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split("\n")
    L = tdl._lib
    assert list(map(int, out[0].split())) == [C.sizeof(L.PhotoArgs), C.sizeof(L.FeatArgs), C.sizeof(L.EdgeArgs),
                                              C.sizeof(L.KernelTime)]
    P = L.PhotoArgs
    assert list(map(int, out[1].split())) == [P.min_depth.offset, P.target.offset, P.noise.offset, P.losses.offset,
                                              P.dP.offset]
    F = L.FeatArgs
    assert list(map(int, out[2].split())) == [F.tgt.offset, F.loss.offset, F.dP.offset, F.layout.offset, F.dtype.offset]
    E = L.EdgeArgs
    assert list(map(int, out[3].split())) == [E.feature.offset, E.d_feature.offset]
    Rc = L.ReconArgs
    assert list(map(int, out[4].split())) == [C.sizeof(Rc), Rc.pred.offset, Rc.d_pred.offset]
    Po = L.PoseArgs
    assert list(map(int, out[5].split())) == [C.sizeof(Po), Po.axisangle.offset, Po.dT.offset, F.bwd_scratch.offset]
    In = L.InputArgs
    assert list(map(int, out[6].split())) == [C.sizeof(In), In.frames.offset, In.jitter.offset, In.color_aug.offset,
                                              In.workspace_bytes.offset]


def test_argument_errors_without_a_gpu(lib, tdl):
    L = tdl._lib
    assert lib.tdl_photo_fwd(None, None) == -1                      # TDL_ERR_NULL
    a = L.PhotoArgs()
    a.B, a.H, a.W, a.S, a.nscales = 1, 64, 96, 5, 4
    assert lib.tdl_photo_fwd(C.byref(a), None) == -4                # TDL_ERR_COUNT (S > TDL_MAX_SRC)
    a.S = 2
    assert lib.tdl_photo_fwd(C.byref(a), None) == -1                # required pointers missing
    dummy = C.create_string_buffer(64)
    p = C.addressof(dummy)
    a.target = a.P = a.invK = a.workspace = a.losses = p
    a.src[0] = a.src[1] = p
    for s in range(4):
        a.disp[s] = p
        a.disp_h[s], a.disp_w[s] = 64 >> (s + 1), 96 >> (s + 1)
    a.disp_h[1] = 15                                                 # 64/15 is not a power of two
    assert lib.tdl_photo_fwd(C.byref(a), None) == -2                # TDL_ERR_SHAPE
    a.disp_h[1] = 16
    a.workspace_bytes = 8
    assert lib.tdl_photo_fwd(C.byref(a), None) == -3                # TDL_ERR_WORKSPACE
    dh = (C.c_int32 * 4)(32, 16, 8, 4)
    dw = (C.c_int32 * 4)(48, 24, 12, 6)
    need = lib.tdl_photo_ws_bytes(1, 64, 96, 2, 4, dh, dw)
    assert need >= 4 * 64 * 96 + 4 * 4 * 8                           # argmin masks + accumulators at least
    assert lib.tdl_feat_fwd(None, None) == -1
    assert lib.tdl_edge_smooth_fwd(None, None) == -1
    assert lib.tdl_pose_fwd(None, None) == -1 and lib.tdl_pose_bwd(None, None) == -1
    pa = L.PoseArgs()
    pa.B, pa.axisangle, pa.translation, pa.T = 0, p, p, p
    assert lib.tdl_pose_fwd(C.byref(pa), None) == -2                 # TDL_ERR_SHAPE
    assert lib.tdl_launch_count(b"tdl_photo_fwd") == 4
    assert lib.tdl_input_fwd(None, None) == -1
    ia = L.InputArgs()
    ia.B, ia.H, ia.W, ia.nframes = 1, 8, 8, 6
    assert lib.tdl_input_fwd(C.byref(ia), None) == -4                # TDL_ERR_COUNT (nframes > TDL_MAX_SRC + 1)
    ia.nframes = 1
    assert lib.tdl_input_fwd(C.byref(ia), None) == -1                # frames[0] missing
    ia.frames[0] = p
    ia.jitter = ia.order = ia.do_aug = ia.workspace = p
    ia.workspace_bytes = 4
    assert lib.tdl_input_fwd(C.byref(ia), None) == -3                # TDL_ERR_WORKSPACE
    assert lib.tdl_input_ws_bytes(8, 3) >= 8 * 3 * 8
    # valid arguments but no sm_100 device behind the call: TDL_ERR_NODEVICE, never a raw cudaError / a launch
    import torch
    if not torch.cuda.is_available():
        pa.B = 1
        assert lib.tdl_pose_fwd(C.byref(pa), None) == -5
        assert b"sm_100" in lib.tdl_strerror(-5)
    # option switches (replace the per-call getenv of ABI v2)
    v = C.c_int(-1)
    assert lib.tdl_get_option(b"photo_sparse_max", C.byref(v)) == 0 and v.value == 128
    assert lib.tdl_set_option(b"photo_sparse_max", 16) == 0
    assert lib.tdl_get_option(b"photo_sparse_max", C.byref(v)) == 0 and v.value == 16
    assert lib.tdl_set_option(b"photo_sparse_max", 128) == 0
    assert lib.tdl_set_option(b"no_such_option", 1) == -6 and lib.tdl_set_option(None, 1) == -1
    with L.options(no_tma=1):
        assert L.get_option("no_tma") == 1
    assert L.get_option("no_tma") == 0
    # backward scratch of the bucketed feature gather (ABI v2): G (chunk x h*w x C floats) + buckets + overflow list, bounded
    # by the batch chunk that keeps G L2-resident (64 MB), so it stops growing with the batch
    assert lib.tdl_feat_bwd_scratch_bytes(0, 64, 96, 320, 2) == 0
    assert lib.tdl_feat_bwd_scratch_bytes(1, 64, 96, 320, 5) == 0
    one = lib.tdl_feat_bwd_scratch_bytes(1, 64, 96, 320, 2)
    assert one >= 96 * 320 * (64 * 4 + 2 * 8 * 8 + 2 * 4)
    eight, sixteen = lib.tdl_feat_bwd_scratch_bytes(8, 64, 96, 320, 2), lib.tdl_feat_bwd_scratch_bytes(16, 64, 96, 320, 2)
    assert one < eight and sixteen == eight


def test_ops_refuse_cpu_tensors(tdl):
    import torch
    cfg = tdl.ops.PhotoConfig(n_src=1, n_scales=1, photo_coef=(1.0,), smooth_coef=(1.0,))
    t = torch.zeros(1, 3, 32, 32)
    with pytest.raises(tdl._lib.TdlError, match="no CPU implementation"):
        tdl.ops.PhotometricSmoothLoss.apply(cfg, t, torch.zeros(1, 1, 3, 4), torch.zeros(1, 3, 3), t,
                                            torch.zeros(1, 1, 16, 16))
