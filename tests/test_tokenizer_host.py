"""Host logic of the Tokenizer mirror: JSON v2.0, special-token splitter goldens, lookups, decode.
CPU only (no encode calls)."""
import json
import pickle

import pytest

import tokengeex
from tokengeex_b200.tokenizer import Tokenizer, TokenGeeXError, fmt_f64, split_special_tokens


# /root/reference/src/tokenizer.rs:442-486  test_special_tokens_splitter
@pytest.mark.parametrize("text,expected,specials", [
    ("<EOS>Hello<EOS>", [("<EOS>", True), ("Hello", False), ("<EOS>", True)], ["<EOS>", "random", "<EOS_2>"]),
    ("randomstring", [("random", True), ("string", False)], ["<EOS>", "random", "<EOS_2>"]),
    ("random<EOS_2>string", [("random", True), ("<EOS_2>", True), ("string", False)], ["<EOS>", "random", "<EOS_2>"]),
    ("nospecialtokens", [("nospecialtokens", False)], ["<EOS>", "random", "<EOS_2>"]),
    ("No special tokens", [("No special tokens", False)], []),
])
def test_ref_special_tokens_splitter(text, expected, specials):
    assert split_special_tokens(text, specials) == expected


def test_splitter_list_order_beats_length():  # Q17
    assert split_special_tokens("a<EOS_2>b", ["<EOS", "<EOS_2>"]) == [("a", False), ("<EOS", True), ("_2>b", False)]
    assert split_special_tokens("", ["x"]) == []
    assert split_special_tokens("你<s>好", ["<s>"]) == [("你", False), ("<s>", True), ("好", False)]


def test_module_surface_matches_pyi():
    assert tokengeex.Tokenizer is Tokenizer and issubclass(tokengeex.TokenGeeXError, Exception)
    for name in ["encode", "encode_ordinary", "encode_batch", "encode_ordinary_batch", "decode", "decode_batch",
                 "token_to_id", "base_token_to_id", "special_token_to_id", "id_to_token", "id_to_base_token",
                 "id_to_special_token", "add_special_tokens", "special_tokens", "is_special", "is_base", "vocab_size",
                 "base_vocab_size", "special_vocab_size", "save", "common_prefix_search", "from_file", "from_str",
                 "to_string", "__getstate__", "__setstate__"]:
        assert hasattr(Tokenizer, name), name


JSON = json.dumps({
    "version": "2.0", "special_tokens": ["<|eos|>", "<|pad|>"],
    "processors": [{"type": "crlf"}, {"type": "unicode", "form": "nfc"}],
    "vocab": [{"value": "def", "score": -7.25}, {"value": "/w", "score": -9.1, "encoded": True},
              {"value": "\n", "score": -3.2, "keep": True}, {"value": "d", "score": -4.0},
              {"value": "你好", "score": -5.5}]})


def test_json_roundtrip_and_lookups(tmp_path):
    t = Tokenizer.from_str(JSON)
    assert t.base_vocab_size() == 5 and t.special_vocab_size() == 2 and t.vocab_size() == 7
    assert t.id_to_base_token(1) == (b"\xff", -9.1)  # "/w" is base64 (no pad) of 0xFF
    assert t.id_to_token(5) == b"<|eos|>" and t.id_to_special_token(6) == "<|pad|>" and t.id_to_token(7) is None
    assert t.token_to_id(b"def") == 0 and t.token_to_id(b"<|pad|>") == 6 and t.token_to_id(b"nope") is None
    assert t.special_token_to_id("<|eos|>") == 5 and t.base_token_to_id(b"\xff") == 1
    assert t.is_special(5) and not t.is_special(4) and not t.is_special(7) and t.is_base(4)
    s = t.to_string()
    assert s.startswith('{"version":"2.0","special_tokens":["<|eos|>","<|pad|>"],"processors":[{"type":"crlf"},'
                        '{"type":"unicode","form":"nfc"}],"vocab":[{"value":"def","score":-7.25},'
                        '{"value":"/w","score":-9.1,"encoded":true},{"value":"\\n","score":-3.2,"keep":true}')
    t2 = Tokenizer.from_str(s)
    assert t2.to_string() == s
    p = tmp_path / "tok.json"
    t.save(str(p))
    t3 = Tokenizer.from_file(str(p))
    assert t3.to_string() == s
    assert json.loads(p.read_text())["vocab"][1] == {"value": "/w", "score": -9.1, "encoded": True}
    assert pickle.loads(pickle.dumps(t)).to_string() == s
    t.add_special_tokens(["<|eos|>", "<|new|>"])  # duplicates ignored (src/tokenizer.rs:44-47)
    assert t.special_tokens() == ["<|eos|>", "<|pad|>", "<|new|>"]
    assert list(t.common_prefix_search("def x")) == [3, 0]


def test_json_errors():
    for bad, msg in [('{"version":"1.0"}', "unsupported version: 1.0"), ('{"vocab":[]}', "missing field `version`"),
                     ('{"version":"2.0","x":1}', "unknown field `x`"),
                     ('{"version":"2.0","vocab":[{"value":"a"}]}', "missing field `score`"),
                     ('{"version":"2.0","vocab":[{"value":"a","score":1,"z":2}]}', "unknown field `z`"),
                     ('{"version":"2.0","processors":[{"type":"unicode","form":"xx"}]}', "untagged enum")]:
        with pytest.raises(TokenGeeXError) as ei:
            Tokenizer.from_str(bad)
        assert msg in str(ei.value)
    # Q18: an object without "type" deserialises as Crlf
    t = Tokenizer.from_str('{"version":"2.0","processors":[{"form":"nfc"}]}')
    assert t.to_string() == '{"version":"2.0","special_tokens":[],"processors":[{"type":"crlf"}],"vocab":[]}'
    with pytest.raises(TokenGeeXError):
        Tokenizer.from_file("/nonexistent/tok.json")


def test_decode():
    t = Tokenizer.from_str(JSON)
    assert t.decode([0, 2, 4, 5, 3, 6], True) == "def\n你好<|eos|>d<|pad|>"
    assert t.decode([0, 2, 4, 5, 3, 6], False) == "def\n你好d"
    assert t.decode([1, 3], True) == "�d"  # from_utf8_lossy
    assert t.decode_batch([[0], [3, 3]], True) == ["def", "dd"]
    with pytest.raises(TokenGeeXError) as ei:
        t.decode([99], True)
    assert str(ei.value) == "token id 99 is out of bounds"


def test_fmt_f64_like_ryu():
    cases = {1.0: "1.0", -7.25: "-7.25", 0.1: "0.1", 1e16: "1e16", 1e15: "1000000000000000.0", 1e-7: "1e-7",
             1.5e-7: "1.5e-7", 0.00001: "0.00001", 0.000015: "0.000015", 123456789.0: "123456789.0",
             1.2345678901234568e17: "1.2345678901234568e17", -6.478377202549375: "-6.478377202549375",
             5e-324: "5e-324", 1.7976931348623157e308: "1.7976931348623157e308", 0.0: "0.0", 100.0: "100.0",
             12345678.9: "12345678.9", 0.001234: "0.001234", 1234e-8: "0.00001234", 1234e-9: "1.234e-6"}
    for x, want in cases.items():
        assert fmt_f64(x) == want, (x, fmt_f64(x), want)
        assert float(fmt_f64(x)) == x


def test_encode_without_device_raises_cleanly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without CUDA")
    t = Tokenizer.from_str(JSON)
    with pytest.raises(TokenGeeXError):
        t.encode("def", 0.0)
    with pytest.raises(TokenGeeXError):
        t.encode("def", 0.5)


def test_add_base_tokens():
    """src/tokenizer.rs:56-61 -> src/model.rs:184-194: new ids at the end, duplicates take the new id."""
    from tokengeex_b200.tokenizer import Tokenizer
    t = Tokenizer([b"a", b"b", b"ab"], [-1.0, -1.0, -1.5], special_tokens=["<s>"], device=None)
    assert t.special_token_to_id("<s>") == 3
    t.add_base_tokens([(b"ba", -1.25), (b"a", -0.5)])
    assert t.base_vocab_size() == 5 and t.vocab_size() == 6
    assert t.base_token_to_id(b"ba") == 3 and t.base_token_to_id(b"a") == 4  # HashMap::insert: the last id wins
    assert t.id_to_base_token(4) == (b"a", -0.5) and t.special_token_to_id("<s>") == 5
    assert list(t.common_prefix_search("ab")) == [4, 2]


def test_id_rows_both_forms():
    """The rows of `encode_batch` (src/tokenizer.rs:93-123) are the same lists whether they are built per row (long
    rows) or sliced from one conversion (short rows); empty rows stay empty."""
    import numpy as np
    from tokengeex_b200.tokenizer import id_rows
    rng = np.random.default_rng(5)
    for T, S in ((4000, 7), (50, 40), (0, 3), (9, 1), (64, 8)):
        ids = rng.integers(0, 1 << 32, T, dtype=np.uint64).astype(np.uint32)
        cut = np.sort(rng.integers(0, T + 1, S - 1)) if S > 1 else np.zeros(0, np.int64)
        id_off = np.concatenate([[0], cut, [T]]).astype(np.uint64)
        rows = id_rows(ids, id_off)
        assert len(rows) == S and all(type(r) is list for r in rows)
        assert [x for r in rows for x in r] == [int(x) for x in ids]
        assert [len(r) for r in rows] == [int(b - a) for a, b in zip(id_off[:-1], id_off[1:])]
        assert all(type(x) is int for r in rows for x in r)
