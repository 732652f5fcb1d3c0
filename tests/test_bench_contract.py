"""bench.py host-side contract (no GPU): defaults, workload table, and the keys of the reference-arm JSON line
(the CPU step itself is stubbed out here; it is exercised for real by `bench.py --impl reference`)."""
import io
import json
import os
import sys
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_defaults_and_workloads(monkeypatch):
    import bench
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = bench.parse()
    assert (a.gpus, a.steps, a.warmup, a.impl, a.workload) == (1, 10, 3, "lvd_b200", "lvd_zeroscope")
    assert set(bench.WORKLOADS) == {"lvd_zeroscope", "lvd_gligen_zeroscope", "lvd_gligen_zeroscope_guided",
                                    "lvd_plus_modelscope256"}
    z = bench.WORKLOADS["lvd_zeroscope"]
    assert (z["frames"], z["h"], z["w"], z["cfg_tflop"], z["guid_tflop"]) == (24, 40, 72, 42.79, 32.59)
    m = bench.WORKLOADS["lvd_plus_modelscope256"]
    assert (m["frames"], m["h"], m["w"], m["gated"], m["guided"]) == (16, 32, 32, True, True)
    assert "576x320" in bench._workload("lvd_zeroscope") and "i % 5 == 0" in bench._workload("lvd_zeroscope")
    assert "GLIGEN" in bench._workload("lvd_gligen_zeroscope")
    assert bench.METRIC.startswith("denoising-step frames/s") and bench.UNIT == "frames/s"
    boxes, pos = bench.bear_layout(24)
    assert len(boxes) == 1 and len(boxes[0]) == 24 and pos == [[2]]
    assert boxes[0][0][0] == 0.0 and abs(boxes[0][-1][2] - (0.83 * 0.8 + 0.17)) < 1e-9


def test_reference_arm_json_line(monkeypatch):
    import bench
    seen = {}

    def fake(sample_frames, pattern, warmup, budget_s=240.0):
        seen.update(pattern=list(pattern), warmup=warmup)
        return dict(frames_per_s=0.5, sec_per_step=4.0, cores=16, sec_guided=6.0, sec_unguided=3.5, sample_frames=8,
                    steps=len(pattern), warmup=warmup)

    monkeypatch.setattr(bench, "cpu_reference_step", fake)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--steps", "2", "--warmup", "1"])
    monkeypatch.setenv("RANK", "0")
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.main()
    line = json.loads(buf.getvalue().strip())
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == "frames/s"
    assert line["value"] == 0.5 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 16
    assert line["e2e"] == {"value": 0.5, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"] == bench._workload("lvd_zeroscope")
    # --steps / --warmup are honoured exactly and reported as executed; the step mix is guided iff i % 5 == 0
    assert seen == {"pattern": [True, False], "warmup": 1} and line["steps"] == 2 and line["warmup"] == 1
    assert line["config"]["sample_frames"] == 8 and "first 8 of the 24 frames" in line["cpu_baseline"]["sample"]
    # ranks other than 0 print nothing
    monkeypatch.setenv("RANK", "1")
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.main()
    assert buf.getvalue() == ""


def test_tile_width_chooser_follows_tile_rounds():
    """ops.pick_block_n (host logic): 256-wide tiles where they fill the persistent grid, 160-wide ones where 256 would
    leave the last round half empty (M = 4320 / 2160: measured 13-16 % faster, tools/perf_gemm_blockn.py; since the
    role loops moved to uniform registers also M = 8640: conv3x3 175 vs 184 us); fixed widths where the epilogue dictates
    them."""
    from lvd_b200 import ops
    t = lambda m: -(-m // 128)  # noqa: E731
    assert ops.pick_block_n(1280, m_tiles=t(138240), k_contract=1280) == 256
    assert ops.pick_block_n(1280, m_tiles=t(34560), k_contract=11520) == 256
    assert ops.pick_block_n(1280, m_tiles=t(8640), k_contract=11520) == 160
    assert ops.pick_block_n(1280, m_tiles=t(4320), k_contract=11520) == 160
    assert ops.pick_block_n(1280, m_tiles=t(2160), k_contract=3840) == 160
    assert ops.pick_block_n(1280, m_tiles=t(1080), k_contract=1280) in (128, 160)
    assert ops.pick_block_n(320, m_tiles=t(138240), k_contract=320) == 160      # 320 = 2 x 160
    assert ops.pick_block_n(640, m_tiles=t(34560), k_contract=640) == 160
    assert ops.pick_block_n(2560, act=ops.ACT_GEGLU, m_tiles=t(138240), k_contract=320) == 256  # value | gate halves
    assert ops.pick_block_n(1280, mn_major=True, m_tiles=t(138240), k_contract=1280) == 256
    assert ops.pick_block_n(1280) == 256 and ops.pick_block_n(320) == 160        # no shape information: widest divisor


def test_fp32_bias_copy_follows_the_bias_tensor():
    """ops._bias_f32 (host logic): the fp32 copy handed to lvd_gemm's specialised epilogues is converted once per bias
    tensor object, follows in-place updates (load_state_dict into existing weights) and is not shared between
    objects -- it must live as long as the bias, because captured CUDA graphs hold its address."""
    import torch
    from lvd_b200 import ops
    b = (torch.arange(16, dtype=torch.float32) / 7).half()
    p1 = ops._bias_f32(b)
    assert ops._bias_f32(b) == p1
    copy = b._lvd_f32[0]
    assert copy.dtype == torch.float32 and torch.equal(copy, b.float())
    b.mul_(2)  # in-place update of the weight: a fresh copy
    ops._bias_f32(b)
    assert torch.equal(b._lvd_f32[0], b.float())
    c = b.clone()
    assert ops._bias_f32(c) != ops._bias_f32(b)
