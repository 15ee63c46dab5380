"""RLP envelope for a shielded withdrawal: known-answer vectors of the RLP spec
(Ethereum yellow paper, appendix B / the ethereum wiki examples) and the round trip through the message shape of
/root/reference/src/types/tx/custom.rs:214-256."""
import json
import os

import pytest

from owshen_b200 import formats as F

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))


@pytest.mark.parametrize("item,hexenc", [
    ("dog", "83646f67"),
    (["cat", "dog"], "c88363617483646f67"),
    ("", "80"),
    ([], "c0"),
    (b"\x00", "00"),
    (b"\x0f", "0f"),
    (b"\x04\x00", "820400"),
    ([[], [[]], [[], [[]]]], "c7c0c1c0c3c0c1c0"),
    ("Lorem ipsum dolor sit amet, consectetur adipisicing elit",
     "b8384c6f72656d20697073756d20646f6c6f722073697420616d65742c20636f6e7365637465747572206164697069736963696e6720656c6974"),
])
def test_rlp_known_answers(item, hexenc):
    enc = F.rlp_encode(item)
    assert enc.hex() == hexenc

    def norm(x):
        return [norm(y) for y in x] if isinstance(x, list) else (x.encode() if isinstance(x, str) else bytes(x))
    assert F.rlp_decode(enc) == norm(item)


def test_rlp_long_string_and_list():
    s = bytes(range(256)) * 4                       # 1024 bytes -> 0xb9 0x04 0x00
    enc = F.rlp_encode(s)
    assert enc[:3] == bytes([0xB9, 0x04, 0x00]) and F.rlp_decode(enc) == s
    lst = [s, b"x"]
    enc = F.rlp_encode(lst)
    assert enc[0] == 0xF9 and F.rlp_decode(enc) == lst


@pytest.mark.parametrize("bad", ["", "8100", "b80100", "83646f", "c883636174", "83646f6700", "b900"])
def test_rlp_rejects_malformed(bad):
    with pytest.raises(ValueError):
        F.rlp_decode(bytes.fromhex(bad))


def test_shielded_withdraw_roundtrip():
    g = GOLD["groth16"]
    proof = bytes.fromhex(g["proof"])
    pub = b"".join(int(x).to_bytes(32, "little") for x in g["public"])
    msg = F.shielded_withdraw_to_rlp(proof, pub)
    assert msg[0] == 0xF9                               # a list longer than 255 bytes
    items = F.rlp_decode(msg)
    assert items[0] == b"shielded-withdraw" and items[1] == proof and len(items) == 5
    assert F.shielded_withdraw_from_rlp(msg) == (proof, pub)
    for broken in (F.rlp_encode(["mint", proof, pub[:32], pub[32:64], pub[64:]]),
                   F.rlp_encode([F.SHIELDED_WITHDRAW_KIND, proof[:-1], pub[:32], pub[32:64], pub[64:]]),
                   F.rlp_encode([F.SHIELDED_WITHDRAW_KIND, proof, pub[:32], pub[32:64]]), msg + b"\0"):
        with pytest.raises(ValueError):
            F.shielded_withdraw_from_rlp(broken)
    with pytest.raises(ValueError):
        F.shielded_withdraw_to_rlp(proof[:-1], pub)


def test_shielded_withdraw_rlp_kat_shared_with_rust():
    """The known answer bindings/rust/route.rs asserts in its own (uncompiled) unit test: both sides must produce these
    bytes, so the constants are read out of the Rust source rather than repeated here."""
    import hashlib
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(__file__)), "bindings", "rust", "route.rs")).read()
    kat_len = int(re.search(r"KAT_LEN: usize = (\d+);", src).group(1))
    prefix = re.search(r'KAT_PREFIX_HEX: &str = "([0-9a-f]+)";', src).group(1)
    digest = re.search(r'KAT_SHA256_HEX: &str = "([0-9a-f]+)";', src).group(1)
    kind = re.search(r'SHIELDED_WITHDRAW_KIND: &str = "([^"]+)";', src).group(1)
    assert kind == F.SHIELDED_WITHDRAW_KIND
    proof = bytes(range(256))
    pub = bytes((7 * i + 3) % 256 for i in range(96))
    msg = F.shielded_withdraw_to_rlp(proof, pub)
    assert len(msg) == kat_len and msg[:24].hex() == prefix and hashlib.sha256(msg).hexdigest() == digest
    assert F.shielded_withdraw_from_rlp(msg) == (proof, pub)
