"""Host-side arithmetic of bench.py: the algorithmic work figures quoted in the roofline objects."""
import importlib.util
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gtb  # noqa: E402,F401
from tinyllama_cpp_b200 import weights as W  # noqa: E402


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_prefill_flops_match_survey_8d():
    b = _bench()
    cfg = W.TINYLLAMA
    fl = b.prefill_flops(cfg, 2048)
    lin = 2 * 968_884_224 * 2048                      # layer linears, SURVEY.md §8(d) config 4
    head = 2 * 65_542_144
    attn = 22 * 32 * sum(2 * 2 * 64 * (i + 1) for i in range(2048))
    assert fl == lin + head + attn
    assert abs(fl - 4.347e12) < 1e9


def test_decode_bytes_match_survey_8d():
    cfg = W.TINYLLAMA
    assert cfg.weight_bytes_per_token(W.Q4) == 581_864_832
    assert cfg.weight_bytes_per_token(W.Q8) == 1_099_078_016
    assert cfg.weight_bytes_per_token(W.F16) == 2_068_852_736
    assert cfg.norm_bytes_per_token() == 184_320
    assert cfg.kv_bytes_per_pos(W.Q4) == 11_968 and cfg.kv_bytes_per_pos(W.F16) == 22_528
    assert cfg.decode_bytes(W.Q4, 1792) == 581_864_832 + 184_320 + 11_968 * 1792


def test_workloads_cover_baseline_configs():
    b = _bench()
    assert set(b.WORKLOADS) == {"q4", "q8", "f16"}
    assert "configs[3]" in b.PREFILL_DESC
