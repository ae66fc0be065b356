"""The reference's wire / on-disk format for DCRTPoly, host side: the library walks archives written by
the reference's own Serial::Serialize and writes archives that are byte-identical to them.  CPU tier (no GPU): the
device upload is one cudaMemcpyAsync per tower from the offsets verified here."""
import numpy as np
import pytest

import ofhe_b200
from oracle_lib import seeded_residues

CASES = [(8, 1, True), (8, 2, True), (16, 3, False), (1 << 10, 5, True), (1 << 12, 7, False), (1 << 14, 2, True)]


@pytest.mark.parametrize("N,L,evaluation", CASES)
def test_scan_reference_archive(reference, N, L, evaluation):
    q, roots = ofhe_b200.dcrt_params(N, L, 60)
    data = seeded_residues(17, q, N)
    blob = reference.dcrt_serialize(q, roots, evaluation, data)
    s = ofhe_b200.cereal_scan_dcrtpoly(blob)
    assert (s["N"], s["n_towers"], s["format"]) == (N, L, 0 if evaluation else 1)
    assert (s["moduli"] == q).all() and (s["roots"] == roots).all()
    for t in range(L):
        tower = np.frombuffer(blob, dtype="<u8", count=N, offset=int(s["offsets"][t]))
        assert (tower == data[t]).all()


@pytest.mark.parametrize("N,L,evaluation", CASES)
def test_write_is_byte_identical_to_reference(reference, N, L, evaluation):
    q, roots = ofhe_b200.dcrt_params(N, L, 60)
    data = seeded_residues(18, q, N)
    assert ofhe_b200.cereal_write_dcrtpoly(q, roots, 0 if evaluation else 1, data) == \
        reference.dcrt_serialize(q, roots, evaluation, data)


def test_write_then_scan_roundtrip_golden_moduli(golden):
    g = golden("ckks_n64")
    q, roots = g["moduli"], g["roots"]  # 40/50/60-bit CKKS chain incl. P: mixed bit lengths in the big-integer product
    data = seeded_residues(19, q, 64)
    blob = ofhe_b200.cereal_write_dcrtpoly(q, roots, 0, data)
    s = ofhe_b200.cereal_scan_dcrtpoly(blob)
    assert (s["moduli"] == q).all() and (s["roots"] == roots).all() and s["N"] == 64
    assert all((np.frombuffer(blob, dtype="<u8", count=64, offset=int(o)) == data[t]).all() for t, o in enumerate(s["offsets"]))


def test_malformed_archives_are_rejected():
    q, roots = ofhe_b200.dcrt_params(16, 2, 60)
    blob = ofhe_b200.cereal_write_dcrtpoly(q, roots, 0, seeded_residues(20, q, 16))
    for bad in (blob[:40], b"\x00" + blob[1:], blob[:5] + b"\xff" * 8 + blob[13:], b""):
        with pytest.raises(ofhe_b200.OfheError):
            ofhe_b200.cereal_scan_dcrtpoly(bad)
    with pytest.raises(ofhe_b200.OfheError):
        ofhe_b200.cereal_scan_dcrtpoly(blob, max_towers=1)
