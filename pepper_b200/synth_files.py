"""Writers for the file formats the readers of pepper_b200/csrc/bamio.cu consume: BAM (BGZF) + BAI, FASTA + FAI.
Written from the SAM/BAM specification (SAMv1 §4.1 BGZF, §4.2 BAM records, §5.2 BAI, §5.3 reg2bin) with Python's zlib;
used by tests and bench scripts to put the seeded synthetic records of synth.py on disk (there is no samtools / pysam in
the image and the reference ships no alignment files)."""
from __future__ import annotations

import struct
import zlib
import numpy as np

from .synth import RecordBatch

BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def reg2bin(beg: int, end: int) -> int:
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def _bgzf_block(payload: bytes, level: int) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    data = co.compress(payload) + co.flush()
    bsize = 18 + len(data) + 8
    assert bsize <= 65536
    return (b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
            + data + struct.pack("<II", zlib.crc32(payload) & 0xffffffff, len(payload)))


def record_ref_len(cigar_words: np.ndarray) -> int:
    ops = cigar_words & 15
    return int((cigar_words[np.isin(ops, [0, 2, 3, 7, 8])] >> 4).sum())


def write_bam(path: str, contigs: list[tuple[str, int]], batches: dict[int, RecordBatch], header_text: str | None = None,
              block_payload: int = 60000, level: int = 1, long_cigar_over: int = 65535, aux: bytes = b"") -> None:
    """`batches`: tid -> coordinate-sorted RecordBatch.  Writes `path` and `path + '.bai'`.  Records are cut into BGZF
    blocks every `block_payload` bytes regardless of record boundaries (records span blocks, as in real files).  CIGARs
    with more than `long_cigar_over` ops use the CG:B,I convention (SAMv1 4.2.2)."""
    if header_text is None:
        header_text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in contigs) + \
            "@RG\tID:rg1\tSM:sample_b200\n"
    head = b"BAM\x01" + struct.pack("<i", len(header_text)) + header_text.encode() + struct.pack("<i", len(contigs))
    for n, l in contigs:
        head += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", l)
    parts = [head]
    u = len(head)
    rec_spans = {}                                      # tid -> list of (beg, end, ustart, uend)
    for tid in sorted(batches):
        b = batches[tid]
        spans = []
        for r in range(b.n_records):
            so, se = int(b.seq_off[r]), int(b.seq_off[r + 1])
            l_seq = se - so
            cig = b.cigar[b.cigar_off[r]:b.cigar_off[r + 1]].astype("<u4")
            rlen = record_ref_len(cig) if cig.shape[0] else 0
            pos = int(b.pos[r])
            end = pos + (rlen if rlen > 0 else 1)
            name = f"r{tid}_{r}".encode() + b"\0"
            codes = _codes(b, so, se)
            if l_seq & 1:
                codes = np.concatenate([codes, np.zeros(1, np.uint8)])
            seq = ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8).tobytes()
            qual = b.qual[so:se].tobytes()
            tags = aux
            cig_field = cig
            if cig.shape[0] > long_cigar_over:
                cig_field = np.array([(l_seq << 4) | 4, (rlen << 4) | 3], dtype="<u4")
                tags = tags + b"CGBI" + struct.pack("<i", cig.shape[0]) + cig.tobytes()
            body = struct.pack("<iiBBHHHiiii", tid, pos, len(name), int(b.mapq[r]), reg2bin(pos, end), cig_field.shape[0],
                               int(b.flag[r]), l_seq, -1, -1, 0) + name + cig_field.tobytes() + seq + qual + tags
            rec = struct.pack("<i", len(body)) + body
            parts.append(rec)
            spans.append((pos, end, u, u + len(rec)))
            u += len(rec)
        rec_spans[tid] = spans
    stream = b"".join(parts)
    # cut into blocks
    ustart, coff, out = [], [], []
    c = 0
    for s in range(0, len(stream), block_payload):
        blk = _bgzf_block(stream[s:s + block_payload], level)
        ustart.append(s)
        coff.append(c)
        out.append(blk)
        c += len(blk)
    eof_coff = c
    out.append(BGZF_EOF)
    with open(path, "wb") as f:
        f.write(b"".join(out))
    ustart_a = np.array(ustart, dtype=np.int64)

    def voff(upos: int) -> int:
        if upos >= len(stream):
            return eof_coff << 16
        k = int(np.searchsorted(ustart_a, upos, side="right") - 1)
        return (coff[k] << 16) | (upos - ustart[k])
    # BAI
    bai = b"BAI\x01" + struct.pack("<i", len(contigs))
    for tid in range(len(contigs)):
        spans = rec_spans.get(tid, [])
        bins: dict[int, list[list[int]]] = {}
        n_win = 0
        linear: dict[int, int] = {}
        for (pos, end, us, ue) in spans:
            vb, ve = voff(us), voff(ue)
            ch = bins.setdefault(reg2bin(pos, end), [])
            if ch and ch[-1][1] == vb:
                ch[-1][1] = ve
            else:
                ch.append([vb, ve])
            for w in range(pos >> 14, ((end - 1) >> 14) + 1):
                if w not in linear or vb < linear[w]:
                    linear[w] = vb
                n_win = max(n_win, w + 1)
        meta = None
        if spans:
            meta = [[voff(spans[0][2]), voff(spans[-1][3])], [len(spans), 0]]        # pseudo-bin 37450 (ignored by readers)
        bai += struct.pack("<i", len(bins) + (1 if meta else 0))
        for bn in sorted(bins):
            bai += struct.pack("<Ii", bn, len(bins[bn]))
            for vb, ve in bins[bn]:
                bai += struct.pack("<QQ", vb, ve)
        if meta:
            bai += struct.pack("<Ii", 37450, 2) + struct.pack("<QQQQ", meta[0][0], meta[0][1], meta[1][0], meta[1][1])
        bai += struct.pack("<i", n_win)
        last = 0
        for w in range(n_win):
            last = linear.get(w, last)
            bai += struct.pack("<Q", last)
    with open(path + ".bai", "wb") as f:
        f.write(bai)


def _codes(b: RecordBatch, so: int, se: int) -> np.ndarray:
    idx = np.arange(so, se)
    byte = b.seq[idx >> 1]
    return np.where(idx & 1, byte & 15, byte >> 4).astype(np.uint8)


def write_fasta(path: str, contigs: list[tuple[str, np.ndarray]], line: int = 60) -> None:
    """FASTA + .fai (name, length, offset, linebases, linewidth)."""
    off = 0
    fai = []
    with open(path, "wb") as f:
        for name, seq in contigs:
            hdr = f">{name} synthetic\n".encode()
            f.write(hdr)
            off += len(hdr)
            s = bytes(np.asarray(seq, dtype=np.uint8))
            fai.append(f"{name}\t{len(s)}\t{off}\t{line}\t{line + 1}\n")
            for i in range(0, len(s), line):
                f.write(s[i:i + line] + b"\n")
            off += len(s) + (len(s) + line - 1) // line
    with open(path + ".fai", "w") as f:
        f.write("".join(fai))
