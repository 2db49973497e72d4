"""CPU tests (-m "not gpu") of the host-side mirror of the reference interface, the BAM
reader/writer, the C-ABI surface and the FFT plan."""

import ctypes
import json
import re
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

from chimeralm_b200 import _lib, tokenizer as T
from chimeralm_b200.bam import BamReader, BamWriter, is_chimeric, parse_bam_file
from chimeralm_b200.callbacks import PredictionWriter, load_predictions_from_folder, resume_read_name
from oracle import tokenizer_oracle as TO

ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden"
BAM = GOLD / "test_chimric_reads.bam"


@pytest.fixture(scope="module")
def tok_gold():
    return json.loads((GOLD / "tokenizer_golden.json").read_text())


def test_character_tokenizer_mirror_matches_reference(tok_gold):
    seqs = tok_gold["seqs"]
    for case in tok_gold["cases"]:
        mml = case["model_max_length"]
        tok = T.CharacterTokenizer(model_max_length=mml, padding_side=case["padding_side"])
        if mml is None:
            got = [tok(s)["input_ids"] for s in seqs]
        else:
            assert tok.max_len_single_sentence == case["max_len_single_sentence"]
            got = [tok(s, truncation=True, max_length=tok.max_len_single_sentence, padding=True)["input_ids"] for s in seqs]
        assert got == case["input_ids"]
        assert tok.pad_token_id == case["pad_token_id"]
        if mml is not None:
            feats = [T.tokenize_and_align_labels_and_quals_ids({"seq": s, "id": f"read-{i}/{len(s)}"}, tok, tok.max_len_single_sentence)
                     for i, s in enumerate(seqs)]
            batch = T.DataCollator(tok).torch_call(feats)
            assert batch["input_ids"].tolist() == case["collated_input_ids"]
            assert batch["id"].tolist() == case["collated_id"] and batch["id"].dtype == torch.int8
            assert batch["labels"].tolist() == case["collated_labels"]


def test_reference_kats_on_mirror():
    tok = T.CharacterTokenizer()
    enc = tok.encode("ATCG")
    assert enc == [0, 7, 10, 8, 9, 1]
    assert tok.convert_ids_to_tokens(enc) == ["[CLS]", "A", "T", "C", "G", "[SEP]"]
    assert tok.decode(enc) == "ATCG"


def test_hub_flavour_tokenizer():
    tok = T.load_tokenizer_from_hyena_model("hyenadna-small-32k-seqlen")
    assert tok.padding_side == "left" and tok.max_len_single_sentence == 32769
    assert tok("ATCGCGTG")["input_ids"] == [7, 10, 8, 9, 8, 9, 10, 9, 1]  # notebooks/attention.ipynb:502: 8 bases + 1 special
    long = tok("A" * 40000, truncation=True, max_length=tok.max_len_single_sentence)["input_ids"]
    assert len(long) == 32769 and long[-1] == 1
    with pytest.raises(ValueError):
        T.load_tokenizer_from_hyena_model("nope")


def test_bam_reader_matches_survey_probe():
    recs = list(parse_bam_file(BAM))
    lens = [len(r["seq"]) for r in recs]
    assert len(recs) == 100 and min(lens) == 524 and max(lens) == 137138 and sum(lens) == 1223444
    assert sum(1 for x in lens if x > 32768) == 11
    assert set("".join(r["seq"][:2000] for r in recs)) <= set("ACGT")
    with BamReader(BAM) as bam:
        flags = {}
        for r in bam:
            assert is_chimeric(r)
            flags[r.flag] = flags.get(r.flag, 0) + 1
    assert flags == {0: 60, 16: 40}


def test_bam_write_roundtrip(tmp_path):
    out = tmp_path / "copy.bam"
    with BamReader(BAM) as bam:
        w = BamWriter(out, bam.header_bytes())
        recs = [r for _, r in zip(range(20), bam)]
        for r in recs:
            w.write(r)
        w.close()
    with BamReader(out) as again:
        back = list(again)
    assert [r.raw for r in back] == [r.raw for r in recs]
    assert out.read_bytes().endswith(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))


def test_prediction_writer_and_loader(tmp_path):
    class Tr:
        global_rank = 3

    names = ["read/1", "m64011_190830_220126/1/ccs", "x"]
    ids = torch.tensor([T.encode_read_name(n) for n in names], dtype=torch.int8)
    logits = torch.tensor([[0.2, 0.1], [0.0, 3.0], [1.0, 1.0]])
    w = PredictionWriter(tmp_path / "pred", "batch")
    w.write_on_batch_end(Tr(), None, (logits, torch.full((3,), -1)), None, {"id": ids}, 7, 0)
    f = tmp_path / "pred" / "3_7.txt"
    assert f.read_text() == "read/1\t0\nm64011_190830_220126/1/ccs\t1\nx\t0\n"
    assert f.read_text().splitlines(keepends=True) == TO.prediction_lines(names, [0, 1, 0])
    assert load_predictions_from_folder(tmp_path / "pred") == {"read/1": 0, "m64011_190830_220126/1/ccs": 1, "x": 0}
    assert [resume_read_name(r) for r in ids] == names
    with pytest.raises(TypeError):
        PredictionWriter(None)  # reference behaviour when -o is omitted


def test_c_abi_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "chimeralm_b200.h").read_text()
    declared = set(re.findall(r"\b(clm_[a-z0-9_]+)\s*\(", header))
    declared -= {"clm_ctx", "clm_status", "clm_dtype", "clm_config"}
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/chimeralm_b200.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    handle = _lib.load()
    assert handle.clm_version().decode().startswith("chimeralm_b200")
    cfg = _lib.clm_config()
    handle.clm_default_config(ctypes.byref(cfg))
    assert (cfg.d_model, cfg.n_layer, cfg.d_inner, cfg.vocab_rows, cfg.max_seq_len) == (256, 4, 1024, 16, 32770)


def test_engine_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from chimeralm_b200.engine import Engine
    from chimeralm_b200.weights import make_state_dict

    with pytest.raises(_lib.ChimeraLMNativeError):
        Engine(make_state_dict(0))


def test_fft_plan_on_host():
    exe = ROOT / "oracle" / "_build" / "fft_host_test"
    if not exe.exists():
        import __graft_entry__ as g

        g.build()
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


def test_weight_factory_is_deterministic_and_ckpt_roundtrip(tmp_path):
    from chimeralm_b200.weights import load_checkpoint, make_state_dict, save_lightning_ckpt

    a, b = make_state_dict(5), make_state_dict(5)
    assert all(torch.equal(a[k], b[k]) for k in a)
    save_lightning_ckpt(a, tmp_path / "m.ckpt")
    c = load_checkpoint(tmp_path / "m.ckpt")
    assert list(c) == list(a) and all(torch.equal(a[k], c[k]) for k in a)


def test_data_module_reference_batching():
    from chimeralm_b200.data import BamDataModule

    tok = T.load_tokenizer_from_hyena_model("hyenadna-small-32k-seqlen")
    dm = BamDataModule(tok, batch_size=12, predict_data_path=BAM)
    dm.setup("predict")
    shapes = [tuple(b["input_ids"].shape) for b in dm.predict_dataloader()]
    # 9 batches (8 x 12 + 4) with these padded lengths
    assert [s[0] for s in shapes] == [12] * 8 + [4]
    assert [s[1] for s in shapes] == [32769, 32769, 22074, 32769, 32769, 32769, 26141, 32769, 31855]
    first = next(iter(dm.predict_dataloader()))
    assert first["input_ids"].dtype == torch.int64 and first["id"].dtype == torch.int8
    assert first["labels"].tolist() == [-1] * 12
    # rank sharding (Lightning's unrepeated sampler): rank r takes samples r, r+W, ...
    dm2 = BamDataModule(tok, batch_size=12, predict_data_path=BAM, rank=1, world_size=2)
    dm2.setup("predict")
    assert dm2.batch_size_per_device == 6
    assert sum(b["input_ids"].shape[0] for b in dm2.predict_dataloader()) == 50


# ------------------------------------------------------------------ native BAM ingest (C++)
def _synth_bam(path, n=300, seed=0, max_len=3000):
    from chimeralm_b200.bam import make_record, minimal_header

    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGTNRYKM=", np.uint8)
    w = BamWriter(path, minimal_header())
    flags = [0, 16, 4, 0x100, 0x800, 0x810, 1, 2048 | 16]
    for i in range(n):
        ln = int(rng.integers(0, max_len)) if i % 17 else (0 if i % 2 else 70000)  # empty and > 1 BGZF block
        seq = alphabet[rng.integers(0, len(alphabet), ln)].tobytes()
        extra = b"XAZsome,text;\0" + b"XBBc" + struct.pack("<i", 3) + b"\1\2\3" if i % 3 == 0 else b"ASi" + struct.pack("<i", 7)
        w.write(make_record(f"r{i}/{'x' * (i % 40)}", seq, flag=flags[i % len(flags)], sa_tag=(i % 4 != 1), extra_aux=extra))
    w.close()


def test_native_ingest_matches_python_reader_on_reference_fixture():
    from chimeralm_b200.bam import parse_bam_file_bytes
    from chimeralm_b200.ingest import read_bam_flat

    ref = list(parse_bam_file_bytes(GOLD / "test_chimric_reads.bam"))
    for kw in ({}, {"block_reads": 7, "block_bytes": 40000}, {"n_threads": 3}):
        names, flat, offs = read_bam_flat(GOLD / "test_chimric_reads.bam",
                                          max_bases=32768, **kw)
        assert names == [n for n, _ in ref]
        for i, (_, s) in enumerate(ref):
            assert np.array_equal(flat[offs[i] : offs[i + 1]], s[:32768])
    names, flat, offs = read_bam_flat(GOLD / "test_chimric_reads.bam", max_bases=100, max_reads=13)
    assert len(names) == 13 and np.all(np.diff(offs) <= 100)
    assert np.array_equal(flat[offs[3] : offs[4]], ref[3][1][:100])


def test_native_ingest_filter_flags_aux_and_block_spanning(tmp_path):
    from chimeralm_b200.bam import parse_bam_file_bytes
    from chimeralm_b200.ingest import read_bam_flat

    p = tmp_path / "synth.bam"
    _synth_bam(p)
    ref = list(parse_bam_file_bytes(p))
    assert 0 < len(ref) < 300
    names, flat, offs = read_bam_flat(p, max_bases=1 << 20, block_reads=5, block_bytes=1 << 17)
    assert names == [n for n, _ in ref]
    assert all(np.array_equal(flat[offs[i] : offs[i + 1]], s) for i, (_, s) in enumerate(ref))
    with BamReader(p) as bam:
        every = [(r.name, r.sequence_bytes()) for r in bam]
    names, flat, offs = read_bam_flat(p, max_bases=1 << 20, chimeric_only=False)
    assert names == [n for n, _ in every] and len(names) == 300
    assert all(np.array_equal(flat[offs[i] : offs[i + 1]], s) for i, (_, s) in enumerate(every))


def test_native_ingest_rejects_damaged_files(tmp_path):
    from chimeralm_b200._lib import ChimeraLMNativeError
    from chimeralm_b200.ingest import read_bam_flat

    p = tmp_path / "synth.bam"
    _synth_bam(p, n=120)
    raw = p.read_bytes()
    with pytest.raises(ValueError, match="cannot open"):
        read_bam_flat(tmp_path / "missing.bam", 100)
    (tmp_path / "text.bam").write_bytes(b"this is not a bam file at all, not even gzip........")
    with pytest.raises(ValueError, match="BGZF"):
        read_bam_flat(tmp_path / "text.bam", 100)
    (tmp_path / "cut.bam").write_bytes(raw[: len(raw) // 2])
    with pytest.raises((ChimeraLMNativeError, ValueError), match="truncated"):
        read_bam_flat(tmp_path / "cut.bam", 100)
    bad = bytearray(raw)
    bad[len(bad) // 2] ^= 0x5A
    (tmp_path / "flip.bam").write_bytes(bytes(bad))
    with pytest.raises((ChimeraLMNativeError, ValueError), match="CRC|inflate|BGZF|record|aux"):
        read_bam_flat(tmp_path / "flip.bam", 100)


def test_datamodule_uses_native_ingest_and_matches_generator(tmp_path):
    from chimeralm_b200.bam import parse_bam_file
    from chimeralm_b200.data import BamDataModule
    from chimeralm_b200.tokenizer import load_tokenizer_from_hyena_model

    tok = load_tokenizer_from_hyena_model("hyenadna-small-32k-seqlen")
    dm = BamDataModule(tok, predict_data_path=GOLD / "test_chimric_reads.bam", batch_size=8, max_predict_samples=20)
    dm.setup("predict")
    ref = list(parse_bam_file(GOLD / "test_chimric_reads.bam"))[:20]
    assert dm.data_predict.names == [r["id"] for r in ref]
    mb = tok.max_len_single_sentence - tok.num_special_tokens
    assert [s.tobytes().decode() for s in dm.data_predict.seqs] == [r["seq"][:mb] for r in ref]
    batches = list(dm.predict_dataloader())
    assert [len(b["id"]) for b in batches] == [8, 8, 4]


def test_streaming_loader_matches_load_all_and_shards_like_the_sampler(tmp_path):
    from chimeralm_b200.data import BamDataModule
    from chimeralm_b200.tokenizer import load_tokenizer_from_hyena_model

    tok = load_tokenizer_from_hyena_model("hyenadna-small-32k-seqlen")
    p = tmp_path / "synth.bam"
    _synth_bam(p, n=200, max_len=400)

    def batches(streaming, rank=0, world=1, limit=None):
        dm = BamDataModule(tok, predict_data_path=p, batch_size=8 * world, max_predict_samples=limit, streaming=streaming,
                           rank=rank, world_size=world)
        dm.setup("predict")
        return list(dm.predict_dataloader())

    for rank, world, limit in ((0, 1, None), (0, 2, None), (1, 2, None), (1, 2, 37), (0, 1, 5)):
        a, b = batches(False, rank, world, limit), batches(True, rank, world, limit)
        assert len(a) == len(b) > 0
        for x, y in zip(a, b):
            assert torch.equal(x["id"], y["id"]) and list(x["indices"]) == list(y["indices"])
            assert torch.equal(x["input_ids"], y["input_ids"]) and torch.equal(x["id"], y["id"])


def test_fastq_and_parquet_predict_inputs(tmp_path):
    """FASTQ (plain / gz, upper-cased like pyfastx(uppercase=True)) and Parquet (columns id, seq) feed the same
    tokenise + collate path as the BAM module; ids are checked against the oracle tokenizer."""
    import gzip

    import pyarrow as pa
    import pyarrow.parquet as pq

    from chimeralm_b200.data import BamDataModule
    from chimeralm_b200.tokenizer import load_tokenizer_from_hyena_model

    tok = load_tokenizer_from_hyena_model("hyenadna-small-32k-seqlen")
    recs = [("r1 extra words", "ACGTNacgtn"), ("r2", "GGGTTTAAACCC"), ("r3/1", "A"), ("r4", "TTTTGGGGCCCCAAAA" * 5)]
    fq = "".join(f"@{n}\n{s}\n+\n{'I' * len(s)}\n" for n, s in recs)
    (tmp_path / "a.fastq").write_text(fq)
    with gzip.open(tmp_path / "b.fq.gz", "wt") as f:
        f.write(fq)
    pq.write_table(pa.table({"id": [n.split()[0] for n, _ in recs], "seq": [s.upper() for _, s in recs],
                             "qual": ["I" * len(s) for _, s in recs]}), tmp_path / "c.parquet")
    want_names = [n.split()[0] for n, _ in recs]
    want_ids = TO.collate([TO.encode(s.upper(), max_length=32769, add_cls=False) for _, s in recs], padding_side="left")
    for path in ("a.fastq", "b.fq.gz", "c.parquet"):
        dm = BamDataModule(tok, predict_data_path=tmp_path / path, batch_size=8, streaming=True)
        dm.setup("predict")
        (batch,) = list(dm.predict_dataloader())
        from chimeralm_b200.callbacks import resume_read_names

        assert resume_read_names(batch["id"]) == want_names, path
        assert batch["input_ids"].tolist() == want_ids, path


def test_native_ingest_block_and_record_carry_across_file_chunks(tmp_path):
    """BGZF blocks and BAM records that straddle the reader's file chunks (default 16 MiB, here a few bytes to a few
    blocks) must come out identical; a synthetic file adds records longer than one BGZF block."""
    from chimeralm_b200.bam import parse_bam_file_bytes
    from chimeralm_b200.ingest import read_bam_flat

    p = tmp_path / "synth.bam"
    _synth_bam(p, n=150, seed=4)
    for path in (GOLD / "test_chimric_reads.bam", p):
        ref = list(parse_bam_file_bytes(path))
        for chunk in (64, 999, 70_000, 200_001):
            names, flat, offs = read_bam_flat(path, 1 << 20, chunk_bytes=chunk, n_threads=3, block_reads=11)
            assert names == [n for n, _ in ref], (path, chunk)
            assert all(np.array_equal(flat[offs[i] : offs[i + 1]], s) for i, (_, s) in enumerate(ref)), (path, chunk)


# ---------------------------------------------------------------------------------- BAI index (`filter`)
def test_bai_matches_the_reference_fixture_index():
    """`tests/golden/test_chimric_reads.bam.bai` is the reference's own samtools-written index of its test BAM
    (`/root/reference/tests/data/`): every bin, chunk, linear-index entry, pseudo-bin and the unplaced-read count agree.
    (Byte equality is not expected: htslib stores a reference's bins in its hash-table order.)"""
    from chimeralm_b200 import bai

    golden = Path(__file__).parent / "golden"
    want = bai.parse_index((golden / "test_chimric_reads.bam.bai").read_bytes())
    got = bai.build_index(golden / "test_chimric_reads.bam")
    assert len(got["refs"]) == len(want["refs"]) == 639
    assert sum(len(r["bins"]) for r in want["refs"]) == 71 and len(want["refs"][0]["linear"]) > 1000
    assert got == want
    assert bai.parse_index(bai.serialize_index(got)) == got


def test_bai_region_queries_equal_a_full_scan():
    """Size-independent property: for any window, reading only the chunks the index names finds exactly the records a
    scan of the whole file finds."""
    import struct as st

    from chimeralm_b200 import bai
    from chimeralm_b200.bam import BamReader

    bam = Path(__file__).parent / "golden" / "test_chimric_reads.bam"
    idx = bai.build_index(bam)
    recs = []
    with BamReader(bam) as f:
        for rec in f:
            l_read_name, n_cigar = rec.raw[8], st.unpack_from("<H", rec.raw, 12)[0]
            recs.append((rec.ref_id, rec.pos, bai._reference_end(rec.raw, rec.pos, rec.flag, l_read_name, n_cigar), rec.name))
    rng = np.random.default_rng(3)
    tids = sorted({r[0] for r in recs if r[0] >= 0})
    for _ in range(60):
        tid = tids[int(rng.integers(len(tids)))]
        on = [r for r in recs if r[0] == tid]
        anchor = on[int(rng.integers(len(on)))]
        beg = max(0, anchor[1] + int(rng.integers(-50_000, 50_000)))
        end = beg + int(rng.integers(1, 200_000))
        want = sorted((r[3], r[1], r[2]) for r in on if r[1] < end and r[2] > beg)
        assert sorted(bai.fetch(bam, idx, tid, beg, end)) == want


def test_bai_small_cases(tmp_path):
    """Edge cases: an empty BAM, unplaced reads at the end (counted, not binned), an unsorted file (refused), reg2bin KATs
    from the SAM specification's bin numbering."""
    from chimeralm_b200 import bai
    from chimeralm_b200.bam import BamWriter, make_record, minimal_header

    assert [bai.reg2bin(0, 1), bai.reg2bin(16383, 16384), bai.reg2bin(16383, 16385), bai.reg2bin(0, 1 << 29)] == [4681, 4681, 585, 0]
    assert bai.reg2bin((1 << 26) - 1, (1 << 26) + 1) == 0 and bai.reg2bin(1 << 26, (1 << 26) + 1) == 4681 + 4096
    hdr = minimal_header((("chr1", 1_000_000), ("chr2", 500_000)))
    empty = tmp_path / "e.bam"
    BamWriter(empty, hdr).close()
    idx = bai.build_index(empty)
    assert idx == {"refs": [{"bins": {}, "linear": []}, {"bins": {}, "linear": []}], "n_no_coor": 0}
    assert bai.parse_index(bai.serialize_index(idx)) == idx

    w = BamWriter(tmp_path / "s.bam", hdr)
    w.write(make_record("a", "ACGT" * 10, ref_id=0, pos=100))
    w.write(make_record("b", "ACGT" * 10, ref_id=0, pos=40_000))
    w.write(make_record("c", "ACGT" * 10, ref_id=1, pos=7))
    w.write(make_record("u1", "ACGT", flag=4, ref_id=-1, pos=-1))
    w.write(make_record("u2", "ACGT", flag=4, ref_id=-1, pos=-1))
    w.close()
    idx = bai.build_index(tmp_path / "s.bam")
    assert idx["n_no_coor"] == 2
    assert idx["refs"][0]["bins"][bai.META_BIN][1] == [2, 0] and idx["refs"][1]["bins"][bai.META_BIN][1] == [1, 0]
    assert len(idx["refs"][0]["linear"]) == 3 and idx["refs"][0]["linear"][1] == idx["refs"][0]["linear"][2]
    assert [h[0] for h in bai.fetch(tmp_path / "s.bam", idx, 0, 39_000, 41_000)] == ["b"]
    assert bai.fetch(tmp_path / "s.bam", idx, 0, 200, 300) == []
    p = bai.index_bam(tmp_path / "s.bam")
    assert p.name == "s.bam.bai" and bai.parse_index(p.read_bytes()) == idx

    w = BamWriter(tmp_path / "u.bam", hdr)
    w.write(make_record("b", "ACGT", ref_id=0, pos=500))
    w.write(make_record("a", "ACGT", ref_id=0, pos=100))
    w.close()
    with pytest.raises(ValueError, match="not coordinate-sorted"):
        bai.build_index(tmp_path / "u.bam")


def test_vectorised_read_name_rows_equal_the_scalar_rule(tok_gold):
    """`encode_read_name_rows` / `resume_read_names` (one numpy pass per batch) against the per-read functions that mirror
    the reference (`tokenizer.py:108-111`, `callbacks.py:38-63`) and against the reference-generated golden rows."""
    from chimeralm_b200.callbacks import resume_read_name, resume_read_names

    names = list(tok_gold["names"]) + ["", "r", "x" * 254, "read/1 with space", "tab\tname", "café"]
    names = [n for n in names if len(n) <= 254]
    want = np.array([T.encode_read_name(n) for n in names], dtype=np.int64).astype(np.int8)
    rows = T.names_to_rows(names)
    got = T.encode_read_name_rows(rows)
    enc = [n.encode("ascii", "replace").decode() for n in names]
    want_enc = np.array([T.encode_read_name(n) for n in enc], dtype=np.int64).astype(np.int8)
    assert got.dtype == np.int8 and np.array_equal(got, want_enc)
    ascii_only = [i for i, n in enumerate(names) if n.isascii()]
    assert np.array_equal(got[ascii_only], want[ascii_only])
    # garbage after the NUL (the native reader does not clear its rows) must not leak
    dirty = rows.copy()
    dirty[:, 200:] = np.where(dirty[:, 200:] == 0, 65, dirty[:, 200:])
    lens = np.array([len(n) for n in enc])
    ok = lens < 199
    assert np.array_equal(T.encode_read_name_rows(dirty)[ok], got[ok])
    back = resume_read_names(torch.from_numpy(got))
    for i, n in enumerate(enc):
        if 0 < len(n) <= 127:
            assert back[i] == resume_read_name(torch.from_numpy(got[i])) == "".join(c for c in n if 32 <= ord(c) <= 126), n
        else:
            # len byte 0, or > 127 and therefore negative after the collator's int8 cast: the scalar function raises (the
            # reference's writer then logs and writes "error_read_i")
            assert isinstance(back[i], ValueError)
            with pytest.raises(ValueError):
                resume_read_name(torch.from_numpy(got[i]))


def test_bucketed_datamodule_deals_token_balanced_batches(tmp_path):
    """`bucket_by_length`: every read exactly once across the ranks, batches under the padded-token budget, ranks balanced by
    cost (LPT) - and the reference's policy (file order, rank::world, fixed size) untouched when the option is off."""
    from chimeralm_b200.data import BamDataModule
    from chimeralm_b200.tokenizer import load_tokenizer_from_hyena_model

    tok = load_tokenizer_from_hyena_model("hyenadna-small-32k-seqlen")
    from chimeralm_b200.bam import make_record, minimal_header

    p = tmp_path / "synth.bam"
    rng = np.random.default_rng(8)
    w = BamWriter(p, minimal_header())
    for i in range(400):
        w.write(make_record(f"read{i}", "ACGT"[i % 4] * int(np.exp(rng.uniform(np.log(100), np.log(3000)))), sa_tag=True))
    w.close()
    for world in (1, 2, 4):
        seen, loads = [], []
        for rank in range(world):
            dm = BamDataModule(tok, predict_data_path=p, batch_size=8 * world, bucket_by_length=True, rank=rank, world_size=world)
            dm.setup("predict")
            lens = dm.data_predict.lengths
            budget = 8 * (int(lens.max()) + 1)
            tokens = 0
            for b in dm.predict_dataloader():
                idx = np.asarray(b["indices"])
                Tb = b["input_ids"].shape[1]
                assert b["input_ids"].shape[0] == len(idx) <= 64 and len(idx) * Tb <= budget
                assert Tb == int(lens[idx].max()) + 1
                seen += idx.tolist()
                tokens += len(idx) * Tb
            loads.append(tokens)
            B, Tm, bud = dm.max_batch_shape()
            assert bud <= budget and Tm == int(lens.max()) + 1 or world > 1
        assert sorted(seen) == list(range(400))            # every read exactly once across the ranks
        assert max(loads) / (sum(loads) / world) < 1.15
    dm = BamDataModule(tok, predict_data_path=p, batch_size=8, rank=1, world_size=2)
    dm.setup("predict")
    first = next(iter(dm.predict_dataloader()))
    assert list(first["indices"]) == [1, 3, 5, 7]
