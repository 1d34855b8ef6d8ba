"""Corpus loader of the `prune` command line (tokengeex_b200/cli.py) against /root/reference/src/cli.rs:237-314."""
import pytest

from tokengeex_b200 import cli
from tokengeex_b200.tokenizer import _Processor


def write(tmp_path, name, samples):
    p = tmp_path / name
    p.write_bytes(b"\x00".join(samples))
    return str(p)


def test_parse_source():
    assert cli.parse_source("code:./a.bin") == ("code", "./a.bin", 1.0)
    assert cli.parse_source("zh:/x/y.bin:0.25") == ("zh", "/x/y.bin", 0.25)
    for bad in ("nopath", "a:b:c:d", "a:b:notanumber"):
        with pytest.raises(ValueError):
            cli.parse_source(bad)


def test_load_sources_semantics(tmp_path):
    raw = [b"def f():\r\n  pass\r\n", b"", "\u4f60\u597d".encode(), b"\r\n", b"x", b"", b"last"]
    path = write(tmp_path, "a.bin", raw)
    # no processors: empty samples dropped, the rest in file order
    (src,) = cli.load_sources([f"a:{path}"])
    assert src.processed_samples == [s for s in raw if s] and src.total_samples == 5
    assert src.total_bytes == sum(len(s) for s in raw)
    # proportion: floor(count * p) non-empty samples FROM THE FRONT, counted before the processors run
    (src,) = cli.load_sources([f"a:{path}:0.5"])
    assert src.processed_samples == [raw[0], raw[2]]
    (src,) = cli.load_sources([f"a:{path}:0.19"])
    assert src.processed_samples == []
    # crlf processor; a sample that becomes empty would be dropped (none does: "\r\n" -> "\n")
    (src,) = cli.load_sources([f"a:{path}"], [_Processor("crlf")])
    assert src.processed_samples == [b"def f():\n  pass\n", "\u4f60\u597d".encode(), b"\n", b"x", b"last"]
    # several sources keep their order
    p2 = write(tmp_path, "b.bin", [b"b1", b"b2"])
    s1, s2 = cli.load_sources([f"a:{path}:0.4", f"b:{p2}"])
    assert [s1.name, s2.name] == ["a", "b"] and s2.processed_samples == [b"b1", b"b2"]


def test_load_sources_rejects_invalid_utf8(tmp_path):
    path = write(tmp_path, "bad.bin", [b"ok", b"\xff\xfe"])
    with pytest.raises(ValueError, match="not valid UTF-8"):
        cli.load_sources([f"bad:{path}"])


def test_prune_cmd_refuses_dropout_outside_unit_interval(tmp_path):
    for bad in (-0.01, 1.0, float("nan")):
        with pytest.raises(ValueError, match="dropout"):
            cli.prune_cmd("in.json", "out.json", 10, ["a:b"], dropout=bad)
