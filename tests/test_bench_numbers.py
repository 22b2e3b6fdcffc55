"""CPU: the algorithmic bytes / FLOPs bench.py divides by are SURVEY.md section 8(d)'s figures (ACE2 1 degree, B = 1)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_algorithmic_figures_match_survey():
    b = _bench()
    assert b.sht_transform_bytes(1) == 384 * (180 * 360 * 4 + 180 * 181 * 8) + 181 * 180 * 180 * 4  # 223.1 MB
    assert abs(b.sht_transform_bytes(1) / 1e6 - 223.1) < 0.1
    a = b.algorithmic(1)
    assert abs(a["dhconv"]["bytes"] / 1e6 - 412.5) < 0.5 and abs(a["dhconv"]["flops"] / 1e9 - 38.4) < 0.1
    assert abs(a["sht.legendre_fwd"]["flops"] / 1e9 - 9.01) < 0.01
    assert abs((a["inner_skip"]["flops"] + a["mlp.fc1"]["flops"] + a["mlp.fc2"]["flops"]) / 1e9 - 95.6) < 0.2
    assert b.STEPS_PER_YEAR == 1460  # 6-hourly steps
    in_names, out_names, prog, forcing, diag = b.names()
    assert (len(in_names), len(out_names), len(prog), len(forcing), len(diag)) == (44, 50, 38, 6, 12)  # ACE2 baseline config
