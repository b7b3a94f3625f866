"""Key / ciphertext blob format (mktfhe_b200/blob.py): byte-exact round trips of the flat upload
layouts, header validation, and (GPU) a context set up from a mapped blob reproducing the committed golden vectors."""
import os

import numpy as np
import pytest

from conftest import keyset
from mktfhe_b200 import blob

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["CGGIparam", "CCS2party"])
def test_key_blob_round_trip(tmp_path, name):
    ks = keyset(name)
    path = str(tmp_path / "keys.blob")
    size = blob.save_keys(path, ks, include_secret=True)
    assert size == os.path.getsize(path) and size % blob.ALIGN == 0
    lk = blob.load_keys(path, verify=True)
    assert lk.params == ks.params and lk.seed == ks.seed
    for a, b in zip(lk.parties, ks.parties):
        for f in ("brk", "ksk", "rlk", "pubb", "lwekey", "ringkey"):
            assert (a[f] is None) == (b.get(f) is None)
            if a[f] is not None:
                assert a[f].dtype == b[f].dtype and np.array_equal(a[f], b[f]), f
    if ks.params.is_mk:
        assert np.array_equal(lk.crs_fft, ks.crs_fft)
    ct = ks.lwe_encrypt_full(1, 77) if ks.params.is_mk else ks.lwe_encrypt(1, 77)
    assert lk.lwe_decrypt(ct) is True and lk.phase(ct) == ks.phase(ct)


def test_evaluation_only_blob_has_no_secrets(tmp_path):
    ks = keyset("CGGIparam")
    path = str(tmp_path / "eval.blob")
    blob.save_keys(path, ks)
    lk = blob.load_keys(path)
    assert lk.lwekeys is None and lk.parties[0]["lwekey"] is None
    with pytest.raises(blob.BlobError, match="evaluation keys only"):
        lk.lwe_decrypt(np.zeros(ks.params.lwe_words, dtype=np.uint32))
    raw = open(path, "rb").read()
    assert ks.parties[0]["lwekey"].tobytes() not in raw
    # the seed regenerates every secret, so an evaluation-only blob must not carry it
    import json, struct
    hlen = struct.unpack("<II", raw[8:16])[1]
    header = json.loads(raw[16:16 + hlen].decode())
    assert header["seed"] is None and lk.seed is None
    assert str(ks.seed).encode() not in raw[:16 + hlen]
    path2 = str(tmp_path / "full.blob")
    blob.save_keys(path2, ks, include_secret=True)
    assert blob.load_keys(path2).seed == ks.seed


def test_ciphertext_blob_and_header_validation(tmp_path):
    ks = keyset("CGGIparam")
    p = ks.params
    cts = np.stack([ks.lwe_encrypt(i & 1, 300 + i) for i in range(5)])
    path = str(tmp_path / "c.blob")
    blob.save_ciphertexts(path, p, cts)
    assert np.array_equal(blob.load_ciphertexts(path, p), cts)
    with pytest.raises(blob.BlobError, match="parameter set"):
        blob.load_ciphertexts(path, keyset("CCS2party").params)
    with pytest.raises(blob.BlobError, match="expected keys"):
        blob.load_keys(path)
    with pytest.raises(ValueError):
        blob.save_ciphertexts(path, p, cts[:, :-1])
    bad = str(tmp_path / "bad.blob")
    open(bad, "wb").write(b"not a blob at all....")
    with pytest.raises(blob.BlobError, match="not a mktfhe-b200 blob"):
        blob.load_ciphertexts(bad)
    raw = bytearray(open(path, "rb").read())
    open(bad, "wb").write(raw[:len(raw) - 4096])
    with pytest.raises(blob.BlobError, match="truncated"):
        blob.load_ciphertexts(bad)
    raw[-3] ^= 0x40                                          # flip a bit of the payload's last page (padding or data)
    raw[len(raw) - 4096 + 5] ^= 0x01
    open(bad, "wb").write(raw)
    with pytest.raises(blob.BlobError, match="checksum"):
        blob.load_ciphertexts(bad)


@pytest.mark.gpu
def test_context_from_a_mapped_blob_reproduces_the_golden_vectors(tmp_path):
    from mktfhe_b200.scheme import MODE_STRICT, setup
    name = "KMS2party"
    ks = keyset(name)
    path = str(tmp_path / "kms2.blob")
    blob.save_keys(path, ks)                                 # evaluation keys only
    lk = blob.load_keys(path, mmap=True)
    s = setup(lk, device=0, mode=MODE_STRICT)
    try:
        g = np.load(os.path.join(GOLDEN, name + ".npz"))
        for op in (0, 3):
            assert np.array_equal(s.gate(op, g["in1"], g["in2"]), g["out"][op])
    finally:
        s.close()
