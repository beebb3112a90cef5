"""CPU, build container only: `patch()` on the REAL reference model.

`create_detector(...)` of /root/reference/src builds the reference's own classes; `patch()` rebinds their forwards to the
functions the mirror classes use.  The kernel front end is replaced by torch-CPU stand-ins (tests/cpu_ops_standin.py), so
everything of the drop-in route except the kernels executes against the reference's attribute names and shapes — and
its outputs are compared with the unpatched model's.  (The kernels are covered by `-m gpu` through the mirror classes,
which share these forward functions; the reference cannot travel to the GPU box.)
"""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
REF_SRC = Path("/root/reference/src")

CODE = r"""
import sys, io, contextlib, copy
sys.path.insert(0, %(root)r); sys.path.insert(0, %(ref)r)
import numpy as np, torch
with contextlib.redirect_stdout(io.StringIO()):
    import encoders, fusion, centernet_target, fusion_detection
import bevfusion_multimodal_3d_object_detection_b200 as b
from bevfusion_multimodal_3d_object_detection_b200 import ops, conv_blocks, synthetic as syn
from tests import cpu_ops_standin

torch.manual_seed(0)
cfg = fusion.load_config(%(cfg)r)
cfg['model']['camera_encoder']['pretrained'] = False
results = {}
for modality in ('lidar+radar', 'all', 'camera_only'):
    with contextlib.redirect_stdout(io.StringIO()):
        model = fusion.create_detector(modality, config=copy.deepcopy(cfg))
    # statistics away from the identity, so that a folding slip shows
    sd = model.state_dict()
    g = torch.Generator().manual_seed(5)
    for k, v in sd.items():
        if k.endswith('running_mean'): sd[k] = torch.randn(v.shape, generator=g) * 0.3
        elif k.endswith('running_var'): sd[k] = torch.rand(v.shape, generator=g) * 1.5 + 0.5
    for k in sd:
        if k.startswith('det_head') and k.endswith('weight'): sd[k] = torch.randn(sd[k].shape, generator=g) * 0.05
    model.load_state_dict(sd)
    model.eval()
    B = 2
    lidar = torch.from_numpy(syn.lidar_batch(11, B, n_valid=900, n_total=1024))
    radars = [torch.from_numpy(r) for r in syn.radar_batch(12, B)]
    imgs = torch.randn(B, 6, 3, 64, 96, generator=g)
    args = (imgs if 'camera' in modality or modality == 'all' else None,
            lidar if 'lidar' in modality or modality == 'all' else None,
            radars if 'radar' in modality or modality == 'all' else None)
    with torch.no_grad():
        ref = model(*args)
        ref_dets = {m.__name__: m.decode_centernet_predictions(ref, score_thresh=0.0, max_detections=40)
                    for m in (centernet_target, fusion_detection)}
    done = b.patch(precision='f32')
    assert {'encoders', 'fusion', 'centernet_target', 'fusion_detection'} <= set(done), done
    saved = {n: getattr(ops, n) for n in cpu_ops_standin.STAND_INS}
    cpu_ops_standin.install(ops)
    try:
        with torch.no_grad():
            got = model(*args)
            got_dets = {m.__name__: m.decode_centernet_predictions(got, score_thresh=0.0, max_detections=40)
                        for m in (centernet_target, fusion_detection)}
        for mod in (model.fusion, model.det_head) + ((model.lidar_encoder,) if args[1] is not None else ()):
            assert mod.b200_precision == 'f32', type(mod).__name__          # patch(precision) reached every module kind
    finally:
        for n, f in saved.items(): setattr(ops, n, f)
        b.unpatch()
    assert set(got) == set(ref)
    for k in ref:
        err = float((got[k] - ref[k]).abs().max() / ref[k].abs().max().clamp_min(1e-30))
        assert err < 1e-5, (modality, k, err)
    for name in ref_dets:
        for d, w in zip(got_dets[name], ref_dets[name]):
            assert set(d) == set(w) and d['labels'].dtype == torch.int64
            assert d['scores'].shape == w['scores'].shape, (modality, name)
            assert torch.allclose(d['scores'], w['scores'], atol=1e-5), (modality, name, float((d['scores'] - w['scores']).abs().max()))
            # same winners in the same order wherever neighbouring scores are further apart than the fp32 noise of the folded BatchNorm
            gap = (w['scores'][:-1] - w['scores'][1:]).abs() > 1e-4
            stable = torch.cat([torch.tensor([True]), gap]) & torch.cat([gap, torch.tensor([True])])
            assert torch.allclose(d['boxes'][stable], w['boxes'][stable], atol=1e-3), (modality, name)
    results[modality] = sorted(ref)
# the patched forwards really went through the kernel front end (not the reference's own layers)
c = cpu_ops_standin.CALLS
assert c['pointnet_encode'] == 2 and c['radar_encode'] == 2 and c['camera_mean'] == 2 and c['lidar_init'] == 2, c
assert c['bilinear_resize'] >= 4 and c['dense_layer'] == 2 and c['centernet_decode'] == 6, c
# ... and through the convolution-stack plumbing (BatchNorm folding, plan cache, the radar branch's 5 x 5 shortcut)
assert c['conv_bn_relu_split'] == 18 and c['border_expand'] == 2 and c['conv_pack_split'] == 18, c
print('PATCHED-REFERENCE-OK', results)
"""


def test_patched_forward_runs_on_the_references_own_detector():
    if not (REF_SRC / "fusion.py").exists():
        pytest.skip("reference checkout not present (GPU box)")
    code = CODE % {"root": str(ROOT), "ref": str(REF_SRC), "cfg": "/root/reference/configs/base.yaml"}
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert "PATCHED-REFERENCE-OK" in r.stdout, r.stdout[-3000:] + r.stderr[-6000:]
