"""CPU (-m "not gpu"): pins the backbone oracle (oracle/darknet_oracle.py) AND the host mirror's variable table
(yolo_tf_b200/model/yolo2/inference.py) to the REFERENCE'S OWN graph builders -- `darknet()`, `_darknet()`, `tiny()`, `_tiny()`
(model/yolo2/inference.py:25-126), `reorg()` and `leaky_relu()` -- executed by tests/golden/make_backbone_golden.py against a
torch float64 stand-in for the slim / tf calls they make (TensorFlow 1.0 is not installable here) ->
tests/golden/backbone_reference.npz.  Pinned: layer sequence, kernel sizes, channel counts, pool placement and stride, the
passthrough tap, reorg's element order, the concat order, variable names + shapes, the center=False variant.  Not pinned:
TF's own conv / BN arithmetic (the stand-in uses the textbook definitions, like the oracle)."""
import os

import numpy as np
import pytest
import torch

from oracle.darknet_oracle import darknet_oracle, init_params, tiny_layer_table, tiny_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "backbone_reference.npz")


def _params(tiny=False):
    return init_params(20, 5, seed=1, table=tiny_layer_table(20, 5) if tiny else None)


@pytest.mark.parametrize("tag,xkey,training", [("darknet", "x64", False), ("darknet_rect", "x96", False), ("darknet_train", "x96", True)])
def test_darknet_oracle_matches_the_reference_graph(tag, xkey, training):
    d = np.load(GOLD)
    taps = {}
    got = darknet_oracle(d[xkey], _params(), 20, 5, training=training, dtype=torch.float64, taps=taps)
    want = d[tag + "_out"]
    assert str(d[tag + "_scope"]) == "yolo2_darknet"
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-11)
    for name in ("conv0", "conv12", "conv19", "conv20"):          # incl. the passthrough source and both concat inputs' consumers
        np.testing.assert_allclose(taps[name][:, :4, :4, :16], d["%s_tap_%s" % (tag, name)], rtol=1e-9, atol=1e-11, err_msg=name)


def test_center_false_variant_is_the_same_arithmetic_with_biases_in_place_of_beta():
    """`_darknet` (inference.py:125-126): BN without beta, then a separate `biases` variable -- with the checkpoint's beta stored
    under `<scope>/conv{i}/biases` the output is identical, which is why the product maps that variable onto beta."""
    d = np.load(GOLD)
    np.testing.assert_allclose(d["darknet_nocenter_out"], d["darknet_out"], rtol=1e-12, atol=1e-14)
    names = [v.split()[0] for v in d["darknet_nocenter_vars"]]
    assert "yolo2_darknet/conv3/biases" in names and not any(n.endswith("/BatchNorm/beta") for n in names)


@pytest.mark.parametrize("tag,xkey", [("tiny", "x96"), ("tiny_nocenter", "x64")])
def test_tiny_oracle_matches_the_reference_graph(tag, xkey):
    d = np.load(GOLD)
    taps = {}
    got = tiny_oracle(d[xkey], _params(tiny=True), 20, 5, dtype=torch.float64, taps=taps)
    assert str(d[tag + "_scope"]) == "yolo2_tiny"
    np.testing.assert_allclose(got, d[tag + "_out"], rtol=1e-9, atol=1e-11)
    for name in ("conv0", "conv5", "conv7"):                      # conv5 feeds the stride-1 SAME pool
        np.testing.assert_allclose(taps[name][:, :4, :4, :16], d["%s_tap_%s" % (tag, name)], rtol=1e-9, atol=1e-11, err_msg=name)


def test_host_mirror_asks_for_exactly_the_variables_the_reference_graph_creates():
    """yolo_tf_b200's `_Engine.sync_weights` looks variables up by the TF names; the table it walks (layer_geometry /
    tiny_layer_geometry) must be the one the reference's graph builders create, name by name and shape by shape."""
    from yolo_tf_b200.model.yolo2 import inference
    d = np.load(GOLD)
    for tag, scope, geom, center in (("darknet", "yolo2_darknet", inference.layer_geometry(20, 5), True),
                                     ("darknet_nocenter", "yolo2_darknet", inference.layer_geometry(20, 5), False),
                                     ("tiny", "yolo2_tiny", inference.tiny_layer_geometry(20, 5), True),
                                     ("tiny_nocenter", "yolo2_tiny", inference.tiny_layer_geometry(20, 5), False)):
        want = {}
        for v in d[tag + "_vars"]:
            name, shp = str(v).split()
            want[name] = tuple(int(s) for s in shp.split("x"))
        mine = {}
        for name, k, cin, cout, has_bn, _ in geom:
            base = "%s/%s" % (scope, name)
            mine[base + "/weights"] = (k, k, cin, cout)
            if has_bn:
                mine[base + "/BatchNorm/gamma"] = (cout,)
                mine[base + ("/BatchNorm/beta" if center else "/biases")] = (cout,)
                mine[base + "/BatchNorm/moving_mean"] = (cout,)
                mine[base + "/BatchNorm/moving_variance"] = (cout,)
            else:
                mine[base + "/biases"] = (cout,)
        assert mine == want, (tag, set(mine) ^ set(want))
    assert inference.darknet.__name__ == "darknet" and inference.tiny.__name__ == "tiny"      # the scope is derived from the name
