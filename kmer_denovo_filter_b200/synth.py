"""Synthetic trio generator for the benchmark configs.

Reference = i.i.d. uniform ACGT (no repeats — say so next to the numbers);
each parent = two haplotypes = reference + SNPs at 1e-3; child = one haplotype
of each parent + heterozygous de novo events (70 % SNV, 15 % insertion, 15 %
deletion, indel length 1-20); paired reads 2 x ``read_len``, insert 400 +- 50,
substitution error 1e-3, N rate 1e-4.  Everything is generated on the GPU with
torch (plumbing: RNG, gathers, bit packing) directly in the packed stream
layout of ``include/kdf.h``; ranks of a multi-GPU run draw disjoint read shards
from the same genomes (same seeds for the genomes, rank-specific read seeds).
"""

import numpy as np

SNP_RATE = 1e-3
ERR_RATE = 1e-3
N_RATE = 1e-4


def _gen(torch, device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def make_reference(torch, device, genome_bp, seed=1000):
    return torch.randint(0, 4, (genome_bp,), dtype=torch.uint8, device=device,
                         generator=_gen(torch, device, seed))


def make_haplotype(torch, ref, seed, rate=SNP_RATE):
    g = _gen(torch, ref.device, seed)
    mask = torch.rand(ref.shape[0], device=ref.device, generator=g) < rate
    shift = torch.randint(1, 4, (ref.shape[0],), dtype=torch.uint8, device=ref.device, generator=g)
    return torch.where(mask, (ref + shift) & 3, ref)


def apply_denovo(torch, hap, n_events, seed=4000):
    """Apply heterozygous de novo events to one haplotype; returns
    ``(new haplotype, events)`` with events = list of (pos, kind, ref, alt)."""
    rng = np.random.RandomState(seed)
    n = hap.shape[0]
    n_snv = int(round(n_events * 0.7))
    n_ins = (n_events - n_snv) // 2
    kinds = ["snv"] * n_snv + ["ins"] * n_ins + ["del"] * (n_events - n_snv - n_ins)
    pos = np.sort(rng.choice(np.arange(1000, n - 1000, 64), size=n_events, replace=False))
    rng.shuffle(kinds)
    pieces = []
    events = []
    cur = 0
    host = None
    for p, kind in zip(pos.tolist(), kinds):
        if kind == "snv":
            old = int(hap[p].item())
            new = (old + int(rng.randint(1, 4))) & 3
            pieces.append(hap[cur:p])
            pieces.append(torch.tensor([new], dtype=torch.uint8, device=hap.device))
            cur = p + 1
            events.append((p, "snv", "ACGT"[old], "ACGT"[new]))
        elif kind == "ins":
            ln = int(rng.randint(1, 21))
            ins = rng.randint(0, 4, size=ln).astype(np.uint8)
            pieces.append(hap[cur:p + 1])
            pieces.append(torch.from_numpy(ins).to(hap.device))
            cur = p + 1
            anchor = "ACGT"[int(hap[p].item())]
            events.append((p, "ins", anchor, anchor + "".join("ACGT"[c] for c in ins.tolist())))
        else:
            ln = int(rng.randint(1, 21))
            pieces.append(hap[cur:p + 1])
            cur = p + 1 + ln
            seg = hap[p:p + 1 + ln].cpu().numpy().tolist()
            events.append((p, "del", "".join("ACGT"[c] for c in seg), "ACGT"[seg[0]]))
    pieces.append(hap[cur:])
    del host
    return torch.cat(pieces), events


def _pack_chunk(torch, bases, valid):
    """(R, L+1) uint8 code / validity matrices → packed int64 / int32 words.
    R*(L+1) must be a multiple of 32."""
    dev = bases.device
    sh_c = (62 - 2 * torch.arange(32, device=dev, dtype=torch.int64))
    sh_v = (31 - torch.arange(32, device=dev, dtype=torch.int64))
    b = bases.reshape(-1, 32).to(torch.int64)
    v = valid.reshape(-1, 32).to(torch.int64)
    codes = ((b * v) << sh_c).sum(dim=1)
    vw = (v << sh_v).sum(dim=1)
    vw = torch.where(vw >= (1 << 31), vw - (1 << 32), vw).to(torch.int32)
    return codes, vw


def make_reads(torch, haps, n_pairs, read_len=150, seed=5000, insert_mean=400.0, insert_sd=50.0,
               chunk_pairs=1 << 19, table=None):
    """Paired reads from a list of haplotypes → packed stream tensors on the device.

    Returns dict(codes int64, valid int32, n_bases, read_starts int64, read_lens int32).
    ``table`` (a dict) additionally receives the reads themselves, for writing them as an
    aligned BAM: ``bases`` uint8 (2 n_pairs, read_len) with N = 4, as sequenced; ``hap``
    (which haplotype), ``start`` / ``insert`` per pair (haplotype coordinates)."""
    dev = haps[0].device
    g = _gen(torch, dev, seed)
    L = read_len
    n_pairs = (n_pairs + 15) // 16 * 16          # 2*n_pairs reads, multiple of 32
    codes_out, valid_out = [], []
    t_bases, t_hap, t_start, t_ins = [], [], [], []
    ar = torch.arange(L, device=dev, dtype=torch.int64)
    done = 0
    while done < n_pairs:
        m = min(chunk_pairs, n_pairs - done)
        which = torch.randint(0, len(haps), (m,), device=dev, generator=g)
        ins = torch.clamp((torch.randn(m, device=dev, generator=g) * insert_sd + insert_mean).round(),
                          L, 4 * insert_mean).to(torch.int64)
        bases = torch.zeros((2 * m, L + 1), dtype=torch.uint8, device=dev)
        starts_all = torch.zeros(m, dtype=torch.int64, device=dev)
        for h, hap in enumerate(haps):
            sel = torch.nonzero(which == h).squeeze(1)
            if sel.numel() == 0:
                continue
            span = hap.shape[0] - ins[sel]
            start = (torch.rand(sel.numel(), device=dev, generator=g, dtype=torch.float64) * span).to(torch.int64)
            starts_all[sel] = start
            r1 = hap[start[:, None] + ar[None, :]]
            r2i = (start + ins[sel] - 1)[:, None] - ar[None, :]
            r2 = 3 - hap[r2i]
            bases[2 * sel, :L] = r1
            bases[2 * sel + 1, :L] = r2
        err = torch.rand((2 * m, L), device=dev, generator=g) < ERR_RATE
        shift = torch.randint(1, 4, (2 * m, L), dtype=torch.uint8, device=dev, generator=g)
        bases[:, :L] = torch.where(err, (bases[:, :L] + shift) & 3, bases[:, :L])
        valid = torch.ones((2 * m, L + 1), dtype=torch.uint8, device=dev)
        valid[:, :L] = (torch.rand((2 * m, L), device=dev, generator=g) >= N_RATE).to(torch.uint8)
        valid[:, L] = 0
        if table is not None:
            t_bases.append(torch.where(valid[:, :L] != 0, bases[:, :L], torch.full_like(bases[:, :L], 4)).cpu())
            t_hap.append(which.cpu())
            t_start.append(starts_all.cpu())
            t_ins.append(ins.cpu())
        c, v = _pack_chunk(torch, bases, valid)
        codes_out.append(c)
        valid_out.append(v)
        done += m
    n_reads = 2 * n_pairs
    n_bases = n_reads * (L + 1) - 1
    if table is not None:
        table.update({"bases": torch.cat(t_bases).numpy(), "hap": torch.cat(t_hap).numpy(),
                      "start": torch.cat(t_start).numpy(), "insert": torch.cat(t_ins).numpy()})
    codes = torch.cat(codes_out)
    valid = torch.cat(valid_out)
    n_words = (n_bases + 31) // 32
    return {"codes": codes[:n_words].contiguous(), "valid": valid[:n_words].contiguous(),
            "n_bases": n_bases,
            "read_starts": torch.arange(n_reads, device=dev, dtype=torch.int64) * (L + 1),
            "read_lens": torch.full((n_reads,), L, dtype=torch.int32, device=dev)}


def pack_sequence_tensor(torch, seq_codes):
    """A single all-valid sequence (uint8 codes on device) → packed stream dict."""
    n = seq_codes.shape[0]
    pad = (-n) % 32
    b = torch.cat([seq_codes, torch.zeros(pad, dtype=torch.uint8, device=seq_codes.device)])
    v = torch.cat([torch.ones(n, dtype=torch.uint8, device=seq_codes.device),
                   torch.zeros(pad, dtype=torch.uint8, device=seq_codes.device)])
    codes_parts, valid_parts = [], []
    step = 1 << 26
    for s in range(0, b.shape[0], step):
        c, vw = _pack_chunk(torch, b[s:s + step], v[s:s + step])
        codes_parts.append(c)
        valid_parts.append(vw)
    return {"codes": torch.cat(codes_parts), "valid": torch.cat(valid_parts), "n_bases": n,
            "read_starts": torch.zeros(1, dtype=torch.int64, device=seq_codes.device),
            "read_lens": torch.tensor([min(n, 2**31 - 1)], dtype=torch.int32, device=seq_codes.device)}


# one packed stream addresses its bases with 32 bits in the PCIe (sparse validity) format
MAX_STREAM_BASES = 1 << 31


def _reads_in_parts(torch, haps, pairs, read_len, seed, max_stream_bases):
    """``make_reads``, as ONE stream dict when the reads fit ``max_stream_bases`` bases,
    else as a list of stream dicts (a whole-genome sample on few GPUs)."""
    per_part = max(16, (max_stream_bases // (2 * (read_len + 1))) // 16 * 16)
    if pairs <= per_part:
        return make_reads(torch, haps, pairs, read_len, seed)
    parts, done, i = [], 0, 0
    while done < pairs:
        m = min(per_part, pairs - done)
        parts.append(make_reads(torch, haps, m, read_len, seed + 100003 * (i + 1)))
        done += m
        i += 1
    return parts


def make_trio(torch, device, genome_bp, depth=30, read_len=150, n_denovo=100, rank=0, world=1,
              seed=1000, max_stream_bases=MAX_STREAM_BASES, keep_genomes=False):
    """Generate the trio; each rank gets 1/world of every sample's reads.

    Returns dict(ref, child, mother, father: packed stream dicts — or lists of them when a
    sample exceeds ``max_stream_bases`` —; events; genome_bp)."""
    ref = make_reference(torch, device, genome_bp, seed)
    m0, m1 = make_haplotype(torch, ref, 2000), make_haplotype(torch, ref, 2001)
    f0, f1 = make_haplotype(torch, ref, 3000), make_haplotype(torch, ref, 3001)
    child_a, events = apply_denovo(torch, m0, n_denovo, 4000)
    child_b = f1
    pairs_total = int(depth * genome_bp / (2 * read_len))
    pairs = pairs_total // world
    out = {"genome_bp": genome_bp, "events": events}
    out["child"] = _reads_in_parts(torch, [child_a, child_b], pairs, read_len, 5000 + 10 * rank,
                                   max_stream_bases)
    out["mother"] = _reads_in_parts(torch, [m0, m1], pairs, read_len, 5001 + 10 * rank, max_stream_bases)
    out["father"] = _reads_in_parts(torch, [f0, f1], pairs, read_len, 5002 + 10 * rank, max_stream_bases)
    # reference shard of this rank, overlapping the next shard by read_len bases
    lo = genome_bp * rank // world
    hi = min(genome_bp, genome_bp * (rank + 1) // world + 64)
    out["ref"] = pack_sequence_tensor(torch, ref[lo:hi])
    out["ref_full_bp"] = genome_bp
    if keep_genomes:     # the BAM writer of the end-to-end benchmark needs them
        out["genomes"] = {"ref": ref, "child": [child_a, child_b], "mother": [m0, m1], "father": [f0, f1]}
    return out
