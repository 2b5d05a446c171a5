"""CPU tests of the host-side logic that needs no device: Blob shape arithmetic, layer
parameters (prototxt defaults), synthetic TREC-QA-shaped data."""
import numpy as np
import pytest

from mms_answer_selection_b200 import synth
from mms_answer_selection_b200.blob import Blob
from mms_answer_selection_b200.layers import LayerParameter, create_layer, CheckError


def test_blob_shape_accessors():
    b = Blob((50, 40, 300), device="cpu")
    assert (b.num(), b.channels(), b.height(), b.width()) == (50, 40, 300, 1)   # sim_cross reads these
    assert b.count() == 600000 and b.count(1) == 12000 and b.count(0, 1) == 50
    b.set_cpu_data(np.arange(600000, dtype=np.float32))
    assert b.cpu_data()[1, 0, 0] == 12000
    b.Reshape((2, 3))
    assert b.data.shape == (2, 3) and not b.data.any()
    with pytest.raises(ValueError):
        Blob((2, 2, 2, 2, 2), device="cpu").num()
    with pytest.raises(TypeError):
        Blob((1,), dtype=np.float16, device="cpu")


def test_layer_parameter_defaults_match_caffe_proto():
    p = LayerParameter("SimCross")
    assert p.sim_cross_param["dist_mode"] == 1 and p.sim_cross_param["mesure_count"] == 1
    assert p.sim_cross_param["bias_term"] is True
    assert p.sim_cross_param["weight_filler"]["type"] == "constant"      # default filler: zeros
    assert LayerParameter("PairRankLoss").pair_rank_loss_param["margin"] == 1.0
    assert LayerParameter("Embed").embed_param["weight_source"] == ""
    with pytest.raises(KeyError):
        LayerParameter("SimCross", sim_cross_param=dict(measure_count=2))  # the reference spells it mesure_count
    with pytest.raises(CheckError, match="Unknown layer type"):
        create_layer(LayerParameter("InnerProduct"))


def test_synthetic_batch_follows_the_reference_padding():
    d = synth.make_qa_batch(N=20, L=40, D=8, mc=2, V=100)
    assert d["idx_q"].dtype == np.float32 and d["idx_q"].shape == (20, 40)
    pad = 99
    for row, lo, hi in ((d["idx_q"], 3, 20), (d["idx_a"], 5, 40)):
        for r in row.astype(int):
            toks = np.flatnonzero(r != pad)
            assert lo <= len(toks) <= hi and r.max() <= pad
            assert toks[0] == (40 - len(toks)) // 2 and np.all(np.diff(toks) == 1)   # centre-padded
    assert np.abs(d["W"]).max() <= 0.08 and not d["B"].any()
    assert synth.pad_sentence(range(50), 40, -1).tolist() == list(range(40))         # truncation


def test_blob_reshape_keeps_storage_when_count_is_unchanged():
    b = Blob((1,), device="cpu")
    b.diff.fill_(2.0)                 # diff allocated before data (SetLossWeights does this)
    b.Reshape((1,))
    assert b.diff.item() == 2.0
    b.Reshape((1, 1))
    assert b.diff.reshape(-1)[0].item() == 2.0
    b.Reshape((3,))
    assert not b.diff.any()


def test_committed_bench_lines_follow_the_contract():
    """The bench lines kept under profiles/ (plain runs on a B200) carry every key the driver's contract names."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for name in ("r01_bench_c2_n1.json", "r01_bench_c3_n4.json", "r02_bench_c3_n1.json", "r02_bench_c3_n8.json"):
        d = json.loads(open(os.path.join(root, "profiles", name)).read().strip().splitlines()[-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
            assert k in d, (name, k)
        assert d["metric"] == "qa_pairs_per_sec_fwd_bwd" and d["higher_is_better"] is True and d["data"] == "synthetic"
        assert "workload" in d["config"] and "model" not in d["config"]
        assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["vs_baseline"] is None
        assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and 0 < d["e2e"]["value"] < d["value"]
        assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
        assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
        assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
        assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
        if d["n_gpus"] == 1:
            assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    ref = json.loads(open(os.path.join(root, "profiles", "r01_bench_ref_c2.json")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["cpu_baseline"]["kind"] in ("reference", "port")
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["value"] == ref["value"]
    # round 2: both arms print the SAME config object (the driver compares them), at every N
    for n in (1, 2, 4, 8):
        ours = json.loads(open(os.path.join(root, "profiles", "r02_bench_c3_n%d.json" % n)).read().strip().splitlines()[-1])
        ref = json.loads(open(os.path.join(root, "profiles", "r02_bench_ref_c3_n%d.json" % n)).read().strip().splitlines()[-1])
        assert ours["config"] == ref["config"] and ours["metric"] == ref["metric"] and ours["unit"] == ref["unit"], n
        assert ours["scaling"] == ref["scaling"] == "strong" and ours["n_gpus"] == ref["n_gpus"] == n
        if n > 1:
            assert ours["parity"]["ok"] is True and ours["parity"]["replicas_bit_identical"] is True


def test_python_mirror_of_the_grouped_scatter_threshold_matches_the_library_source():
    """`_lib.EMBED_GROUPED_MIN_ROWS` duplicates csrc/embed_sorted.cu's kMinRows (MMSNet uses it to decide whether the pair
    call would only run the two per-layer kernels one after the other)."""
    import os
    import re
    from mms_answer_selection_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "mms_answer_selection_b200", "csrc", "embed_sorted.cu")).read()
    m = re.search(r"constexpr long long kMinRows = (\d+);", src)
    assert m and int(m.group(1)) == _lib.EMBED_GROUPED_MIN_ROWS


def test_shipped_library_reads_no_developer_knobs_from_the_environment():
    """Kill switches and probe knobs are compiled in only with -DMMS_DEV_KNOBS / -DMMS_BWD_PROBES: every getenv in the
    product sources is the tracing switch, inside the gated helper, or inside an #ifdef block of those macros."""
    import glob
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "mms_answer_selection_b200", "csrc")
    bad = []
    for path in glob.glob(os.path.join(csrc, "*.cu*")) + glob.glob(os.path.join(csrc, "tc", "*.cu*")):
        depth = 0                                   # nesting inside #ifdef MMS_DEV_KNOBS / MMS_BWD_PROBES blocks
        stack = []
        for n, line in enumerate(open(path), 1):
            t = line.strip()
            if re.match(r"#\s*if", t):
                stack.append(bool(re.match(r"#\s*ifdef\s+(MMS_DEV_KNOBS|MMS_BWD_PROBES)\b", t)))
            elif re.match(r"#\s*else", t) and stack:
                stack[-1] = False
            elif re.match(r"#\s*endif", t) and stack:
                stack.pop()
            gated = any(stack)
            if "getenv(" in line and not gated and "MMS_TC_TRACE" not in line:
                bad.append("%s:%d: %s" % (os.path.relpath(path, root), n, t))
    assert not bad, "\n".join(bad)
