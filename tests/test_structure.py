"""Host-side logic on the CPU (-m "not gpu"): layer API mirror, state-dict manifests, tracing,
fusion and lowering, the drop-in shim under the reference's own model files, weight I/O."""
import collections
import os

import pytest
import torch

import tlxcv_b200 as tlx
from oracle import ref_loader, tlx_compat
from tlxcv_b200 import models, nn, planner, runtime, tasks
from tlxcv_b200.planner import Shape
from tlxcv_b200.testing import seeded_state_dict

ALL = ["resnet18", "resnet34", "resnet50", "resnet101", "wide_resnet50_2", "resnext50_32x4d", "resnext50_64x4d",
       "mobilenet_v1", "mobilenet_v2", "darknet53_cls", "darknet53_det"]


@pytest.mark.parametrize("name", ALL)
def test_state_dict_manifest_equals_reference(name, manifests):
    """Same keys, shapes AND order as the reference's own file => state dicts and positional .npz load unchanged."""
    model = models.REGISTRY[name]()
    got = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    assert got == manifests[name]


def _plan(name, n=2):
    m = models.REGISTRY[name]().set_eval()
    if name == "darknet53_det":
        return planner.plan_for_shapes(m, {"images": Shape(n, 3, 608, 608)})
    return planner.plan_for_shapes(m, Shape(n, 3, 224, 224))


# (conv ops, residual adds fused into a conv epilogue)
EXPECT = {"resnet50": (53, 16), "resnet18": (20, 8), "resnext50_32x4d": (53, 16), "mobilenet_v2": (52, 10),
          "mobilenet_v1": (27, 0), "darknet53_cls": (52, 23), "darknet53_det": (52, 23)}


@pytest.mark.parametrize("name", sorted(EXPECT))
def test_everything_fuses_into_conv_epilogues(name):
    spec, _ = _plan(name)
    kinds = collections.Counter(planner.OP_NAMES[o.kind] for o in spec.ops)
    convs, fused = EXPECT[name]
    assert kinds["conv"] == convs
    assert sum(1 for o in spec.ops if o.kind == planner.OP_CONV and o.in1 >= 0) == fused
    assert kinds["add_act"] == 0, "an add / activation escaped fusion"
    assert all(o.bn is not None for o in spec.ops if o.kind == planner.OP_CONV)
    if name == "darknet53_det":
        assert spec.out_shapes == [(2, 256, 76, 76), (2, 512, 38, 38), (2, 1024, 19, 19)]
        assert kinds["export_nchw"] == 3
    else:
        assert spec.out_shapes == [(2, 1000)] and kinds["linear"] == 1 and kinds["gap"] == 1


def test_resnet_bottleneck_epilogues():
    spec, _ = _plan("resnet50")
    by_path = {o.path: o for o in spec.ops}
    c3, ds = by_path["layer1.0.conv3"], by_path["layer1.0.downsample.0"]
    # conv3+bn3 has no activation of its own; the later downsample conv takes it as residual and applies the ReLU
    assert c3.act1 == planner.ACT_NONE and c3.in1 < 0
    assert ds.in1 == c3.out and ds.act2 == planner.ACT_RELU
    c3b = by_path["layer1.1.conv3"]
    assert c3b.in1 >= 0 and c3b.act2 == planner.ACT_RELU
    stem = by_path["conv1"]
    assert (stem.r, stem.s, stem.stride, stem.pad, stem.act1) == (7, 7, 2, 3, planner.ACT_RELU)
    assert by_path["layer2.0.conv2"].stride == 2      # v1.5: stride on the 3x3


def test_darknet_leaky_then_add():
    spec, _ = _plan("darknet53_det", 1)
    op = next(o for o in spec.ops if o.path.endswith("basicblock0.conv2.conv"))
    assert op.act1 == planner.ACT_LEAKY and abs(op.alpha1 - 0.1) < 1e-7 and op.in1 >= 0 and op.act2 == planner.ACT_NONE


def test_mobilenet_v2_relu6_depthwise_residual():
    spec, _ = _plan("mobilenet_v2")
    dw = [o for o in spec.ops if o.kind == planner.OP_CONV and o.groups > 1]
    assert len(dw) == 17 and all(o.act1 == planner.ACT_RELU6 for o in dw)
    proj = next(o for o in spec.ops if o.path == "features.3.conv.2")
    assert proj.act1 == planner.ACT_NONE and proj.in1 >= 0


def test_predict_is_one_plan_with_argmax():
    m = models.resnet50().set_eval()
    task = tasks.ImageClassification(m)
    spec, _ = planner.plan_for_shapes(task._predictor, Shape(4, 3, 224, 224))
    assert planner.OP_NAMES[spec.ops[-1].kind] == "argmax"
    assert spec.out_shapes == [(4,)] and spec.out_dtypes == [planner.DT_I64]
    assert list(task.state_dict().keys()) == ["backbone." + k for k in m.state_dict().keys()]


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is not mounted here")
@pytest.mark.parametrize("name", ["resnet50", "resnext50_32x4d", "mobilenet_v1", "mobilenet_v2", "darknet53_cls",
                                  "darknet53_det", "yolov3_darknet53", "mobilenet_v1_det", "resnet50_vd", "resnet18_vd",
                                  "resnest50"])
def test_reference_files_run_unmodified_on_the_product_shim(name, manifests):
    """Drop-in: the reference's own model file, imported against tlxcv_b200 installed as `tensorlayerx`,
    builds B200-backed modules with the reference manifest and traces to the same plan as our model."""
    import contextlib

    @contextlib.contextmanager
    def shim():
        import sys

        with tlx_compat.installed():         # supplies the inert paddle / paddle2tlx stubs two files import
            with tlx.as_tensorlayerx():      # ... and the product takes the tensorlayerx names
                # mobilenetv2.py:98 takes its Dropout from paddle2tlx; give it the product's
                sys.modules["paddle2tlx.pd2tlx.ops.tlxops"].tlx_Dropout = nn.Dropout
                yield

    ref_model = ref_loader.build(name, shim=shim)
    assert isinstance(ref_model, nn.Module)
    assert [(k, tuple(v.shape)) for k, v in ref_model.state_dict().items()] == manifests[name]
    ref_model.set_eval()
    ours = models.REGISTRY[name]().set_eval()
    from tlxcv_b200.testing import DICT_INPUT
    arg = {"images": Shape(1, 3, 96, 96)} if name in DICT_INPUT else Shape(1, 3, 96, 96)
    a, _ = planner.plan_for_shapes(ref_model, arg)
    b, _ = planner.plan_for_shapes(ours, arg)
    key = lambda s: [(o.kind, o.in0, o.in1, o.out, o.r, o.s, o.stride, o.pad, o.groups, o.act1, o.act2, o.path)  # noqa: E731
                     for o in s.ops]
    assert key(a) == key(b)
    assert a.out_shapes == b.out_shapes


def test_training_mode_and_cpu_inputs_fail_loudly():
    m = models.resnet18()
    with pytest.raises(NotImplementedError, match="training mode"):
        planner.plan_for_shapes(m, Shape(1, 3, 64, 64))          # is_train is True until set_eval()
    m.set_eval()
    with pytest.raises(runtime.B200RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="no eager path"):
        tlx.add(torch.ones(1), torch.ones(1))


def test_unsupported_patterns_are_rejected():
    class Flat(nn.Module):
        def __init__(self):
            super().__init__()
            self.c = nn.GroupConv2d(in_channels=8, out_channels=8, kernel_size=1, padding=0, b_init=None)

        def forward(self, x):
            return tlx.reshape(self.c(x), (x.shape[0], -1))      # NCHW-order flatten of a 4x4 map

    with pytest.raises(NotImplementedError, match="reshape"):
        planner.plan_for_shapes(Flat().set_eval(), Shape(1, 8, 4, 4))
    with pytest.raises(NotImplementedError, match="channels_first"):
        nn.BatchNorm2d(num_features=4, data_format="channels_last")


def test_weight_files_roundtrip(tmp_path):
    a, b = models.resnet18(), models.resnet18()
    a.load_state_dict(seeded_state_dict(a.state_dict(), "resnet18"))
    p = str(tmp_path / "model.npz")
    a.save_weights(p)                                  # positional list over all_weights (tensorlayerx .npz)
    b.load_weights(p)
    for (k, v), (_, w) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(v, w), k
    p2 = str(tmp_path / "model_dict.npz")
    a.save_weights(p2, format="npz_dict")
    c = models.resnet18().load_weights(p2, format="npz_dict")
    assert torch.equal(c.fc.weights, a.fc.weights)
    # Linear weights stored (out, in) are accepted too
    sd = a.state_dict()
    sd["fc.weights"] = sd["fc.weights"].t().contiguous()
    d = models.resnet18()
    d.load_state_dict(sd)
    assert torch.equal(d.fc.weights, a.fc.weights)


def test_list_held_submodules_register_like_the_oracle():
    m = models.DarkNet()
    keys = list(m.state_dict().keys())
    assert "darknet_conv_block_list.0.basicblock0.conv1.conv.filters" in keys
    assert "downsample_list.3.conv_bn_layer.batch_norm.moving_var" in keys
    assert isinstance(m.darknet_conv_block_list, list) and len(m.darknet_conv_block_list) == 5
    rx = models.resnext50_32x4d()
    assert not any(k.startswith("block_list.") for k in rx.state_dict())   # already registered as bb_i_j


def test_uint8_preprocessing_is_the_plan_input_op():
    """vision.Preprocessed: the uint8 NHWC batch enters through import_u8_nhwc (Normalize + ToTensor fused into the
    layout pass, demo/image_classification/predict-resnet.py:50-56), not through the fp32 NCHW import."""
    from tlxcv_b200 import models, planner, vision

    net = vision.Preprocessed(models.resnet18(), mean=(125.31, 122.95, 113.86), std=(62.99, 62.09, 66.70)).set_eval()
    spec, _ = planner.plan_for_shapes(net, planner.Shape(4, 224, 224, 3, u8=True))
    kinds = [k for k, _ in spec.summary()]
    assert kinds[0] == "import_u8_nhwc" and "import_nchw" not in kinds
    t_in = spec.tensors[spec.inputs[0]]
    assert (t_in.n, t_in.h, t_in.w, t_in.c, t_in.dtype) == (4, 224, 224, 3, planner.DT_U8)
    assert kinds[1] == "conv" and kinds[-1] == "linear"
    import pytest
    with pytest.raises(ValueError):
        vision.NormalizeToTensor(mean=(1.0, 2.0), std=(1.0, 0.0))
