"""The CPU oracle against its pins (-m "not gpu").

* oracle.restated vs the committed golden fixtures (outputs of the REFERENCE'S OWN model files,
  tests/golden/make_golden.py) — everywhere;
* oracle.restated and the fixtures vs oracle.ref_loader (the reference files executed live) —
  only where /root/reference is mounted (the build container).
"""
import os

import numpy as np
import pytest
import torch

from oracle import ref_loader, restated, tlx_compat
from tlxcv_b200.testing import (flatten_outputs, model_input, seeded_state_dict, state_dict_digest, structured_images,
                                synthetic_images)

GOLDEN = ["resnet50", "resnet18", "resnext50_32x4d", "mobilenet_v1", "mobilenet_v2", "darknet53_cls", "darknet53_det",
          "yolov3_darknet53", "mobilenet_v1_det", "resnet50_vd", "resnet18_vd", "resnest50"]
CLASSIFIERS = ["resnet50", "resnet18", "resnext50_32x4d", "mobilenet_v1", "mobilenet_v2", "darknet53_cls", "resnest50"]
# fixtures were minted with oneDNN on the build container's CPU; another CPU may pick other conv kernels
ATOL = 2e-5


def _load(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    outs = [torch.from_numpy(g[k]) for k in sorted(k for k in g.files if k.startswith("out"))]
    return int(g["n"]), int(g["size"]), str(g["weight_digest"]), str(g["input_digest"]), outs


@pytest.mark.parametrize("name", GOLDEN)
def test_restated_matches_golden(name, golden_dir, manifests):
    n, size, wdig, idig, outs = _load(golden_dir, name)
    sd = seeded_state_dict(dict(manifests[name]), name)
    assert state_dict_digest(sd) == wdig, "seeded weight recipe drifted from the one the fixtures were minted with"
    x = synthetic_images(n, size)
    assert state_dict_digest({"x": x}) == idig
    ys = flatten_outputs(restated.forward(name, sd, model_input(name, x)))
    assert len(ys) == len(outs)
    for a, b in zip(ys, outs):
        assert a.shape == b.shape
        assert float((a - b).abs().max()) <= ATOL * max(1.0, float(b.abs().max()))
    if name in CLASSIFIERS:
        assert torch.equal(ys[0].argmax(1), outs[0].argmax(1))


FULL_SIZE = ["resnet50_bs256", "mobilenet_v2_bs64", "resnext50_32x4d_bs64", "darknet53_det_608", "yolov3_darknet53_608"]
SUB = (slice(None), slice(None, None, 4), slice(None, None, 3), slice(None, None, 3))


@pytest.mark.parametrize("fname", FULL_SIZE)
def test_restated_matches_full_size_golden(fname, golden_dir, manifests):
    """The BASELINE.json configurations at their stated size (ResNet-50: the whole bs256 batch; DarkNet-53 at 608x608):
    fixtures minted by the reference's own files on testing.structured_images."""
    g = np.load(os.path.join(golden_dir, f"{fname}.npz"))
    name, n, size, sub = str(g["model"]), int(g["n"]), int(g["size"]), bool(int(g["subsampled"]))
    sd = seeded_state_dict(dict(manifests[name]), name)
    assert state_dict_digest(sd) == str(g["weight_digest"])
    x = structured_images(n, size)
    assert state_dict_digest({"x": x}) == str(g["input_digest"])
    ys = flatten_outputs(restated.forward(name, sd, model_input(name, x)))
    for i, a in enumerate(ys):
        b = torch.from_numpy(g[f"out{i}"])
        scale = max(1.0, float(b.abs().max()))
        if sub:
            assert tuple(a.shape) == tuple(int(v) for v in g[f"shape{i}"])
            assert float((a[SUB] - b).abs().max()) <= ATOL * scale
            assert abs(float(a.double().sum()) - float(g[f"sum{i}"])) <= 1e-6 * float(g[f"abssum{i}"])
            assert abs(float(a.double().abs().sum()) - float(g[f"abssum{i}"])) <= 1e-6 * float(g[f"abssum{i}"])
        else:
            assert float((a - b).abs().max()) <= ATOL * scale
            assert torch.equal(a.argmax(1), b.argmax(1))
            assert len(set(b.argmax(1).tolist())) >= (3 if n >= 256 else 1)      # the images do differ


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is not mounted here")
@pytest.mark.parametrize("name", ["resnet50", "resnext50_32x4d", "mobilenet_v2", "mobilenet_v1", "darknet53_cls",
                                  "darknet53_det", "yolov3_darknet53", "mobilenet_v1_det", "resnet50_vd", "resnet18_vd",
                                  "resnest50"])
def test_restated_matches_reference_files_live(name):
    """Bit-for-bit: the restatement and the reference's own file, same weights, same input."""
    torch.manual_seed(0)
    model = ref_loader.build(name)
    sd = seeded_state_dict(model.state_dict(), name, seed=77)
    model.load_state_dict(sd)
    model.set_eval()
    x = synthetic_images(1, 64 if name in ("darknet53_det", "yolov3_darknet53") else 96, seed=5)
    with torch.no_grad():
        ref = flatten_outputs(model(model_input(name, x)))
    got = flatten_outputs(restated.forward(name, sd, model_input(name, x)))
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert torch.equal(a, b)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is not mounted here")
def test_reference_manifests_match_committed(manifests):
    for name in ("resnet50", "mobilenet_v2", "darknet53_det", "yolov3_darknet53", "mobilenet_v1_det", "resnet50_vd", "resnest50",
                 "resnest101"):
        model = ref_loader.build(name)
        assert [(k, tuple(v.shape)) for k, v in model.state_dict().items()] == manifests[name]


def test_compat_layer_op_semantics():
    """Per-op restatements in the stand-in: BN eval formula, -inf max-pool padding, (in,out) Linear."""
    with tlx_compat.installed() as tlx:
        nn = tlx.nn
        bn = nn.BatchNorm2d(num_features=3, data_format="channels_first")
        with torch.no_grad():
            bn.gamma.copy_(torch.tensor([1.0, 2.0, 0.5]))
            bn.beta.copy_(torch.tensor([0.1, -0.2, 0.0]))
            bn.moving_mean.copy_(torch.tensor([0.5, 0.0, -1.0]))
            bn.moving_var.copy_(torch.tensor([4.0, 1.0, 0.25]))
        bn.set_eval()
        x = torch.randn(2, 3, 4, 4)
        want = (x - bn.moving_mean.view(1, 3, 1, 1)) / torch.sqrt(bn.moving_var.view(1, 3, 1, 1) + 1e-5) \
            * bn.gamma.view(1, 3, 1, 1) + bn.beta.view(1, 3, 1, 1)
        assert torch.allclose(bn(x), want, atol=1e-6)
        mp = nn.MaxPool2d(kernel_size=3, stride=2, padding=1, data_format="channels_first")
        neg = -torch.ones(1, 1, 4, 4)
        assert float(mp(neg).max()) == -1.0          # zero padding would give 0
        lin = nn.Linear(in_features=4, out_features=2)
        assert tuple(lin.weights.shape) == (4, 2)
        conv = nn.GroupConv2d(in_channels=4, out_channels=8, kernel_size=3, padding=1, b_init=(),
                              data_format="channels_first")
        assert conv.biases is None and tuple(conv.filters.shape) == (8, 4, 3, 3)
        assert tlx.argmax(torch.tensor([[0.0, 2.0, 1.0]]), axis=-1).item() == 1


def test_recipe_keeps_logits_small(manifests):
    """App. D: residual-branch damping + FC gain keep logits O(0.3) so 1e-2 absolute is meaningful."""
    sd = seeded_state_dict(dict(manifests["resnet18"]), "resnet18")
    y = restated.forward("resnet18", sd, synthetic_images(2, 96))
    assert 0.05 < float(y.std()) < 1.0
    assert torch.isfinite(y).all()


def test_cv_resize_restatement_against_opencv_vectors_and_live(golden_dir):
    """oracle/cv_resize.py (the Resize of the reference's transform pipeline = cv2 INTER_LINEAR on uint8) against vectors
    minted from cv2, and against cv2 itself where it imports."""
    from oracle.cv_resize import resize_u8

    g = np.load(os.path.join(golden_dir, "cv_resize.npz"))
    k = 0
    while f"src{k}" in g.files:
        src, dst = g[f"src{k}"], g[f"dst{k}"]
        assert np.array_equal(resize_u8(src, dst.shape[0], dst.shape[1]), dst), k
        k += 1
    assert k >= 5
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    for hs, ws, h, w in [(375, 500, 224, 224), (100, 133, 224, 224), (300, 200, 608, 608), (64, 64, 64, 64), (2, 2, 7, 5)]:
        img = rng.integers(0, 256, (hs, ws, 3), dtype=np.uint8)
        assert np.array_equal(resize_u8(img, h, w), cv2.resize(img, (w, h), interpolation=cv2.INTER_LINEAR))
