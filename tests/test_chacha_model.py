"""CPU pin of the CSPRNG model: RFC 8439 known-answer vectors for the block function, sampler sanity."""
import struct

from chacha_model import Model, block

RFC_KEY = bytes(range(32))


def test_rfc8439_block_function_vector():
    # RFC 8439 section 2.3.2: key 00..1f, nonce 00 00 00 09 00 00 00 4a 00 00 00 00, block counter 1
    key = struct.unpack("<8I", RFC_KEY)
    out = block(key, 1, 0x09000000, 0x4A000000, 0x00000000)
    assert out[:4] == [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3]
    assert out[12:] == [0xD19C12B5, 0xB94E16DE, 0xE883D0CB, 0x4E3C50A2]


def test_rfc8439_keystream_vector():
    # RFC 8439 section 2.4.2 (first keystream block of the encryption example): nonce 00 00 00 00 00 00 00 4a 00 00 00 00, counter 1
    key = struct.unpack("<8I", RFC_KEY)
    out = block(key, 1, 0x00000000, 0x4A000000, 0x00000000)
    ks = struct.pack("<16I", *out)
    assert ks[:16].hex() == "224f51f3401bd9e12fde276fb8631ded"


def test_samplers_are_in_range_and_deterministic():
    m1, m2 = Model(RFC_KEY, b"\x01" * 8), Model(RFC_KEY, b"\x01" * 8)
    bound = (1 << 200) + 12345
    a = m1.below(bound, 50, min_one=True)
    assert a == m2.below(bound, 50, min_one=True) and all(1 <= v < bound for v in a)
    assert m1.below(bound, 5) != a[:5]                       # the next call uses another stream
    assert all(v < (1 << 70) for v in m1.bits(70, 20, 3))
    assert set(m1.coins(1000)) == {0, 1}
    for p in m1.permutations(17, 20):
        assert sorted(p) == list(range(17))
