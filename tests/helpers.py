"""Synthetic FASTA / BAM / VCF fixtures in the style of the reference's
tests/helpers.py (which builds them with pysam): an MD5-derived 200-bp
reference, BAMs of a few reads written by the oracle's BAM writer, tiny VCFs."""
import hashlib

from oracle import bam as obam


def create_ref_fasta(path, chrom="chr1", length=200):
    seq = "".join("ACGT"[int(hashlib.md5(str(i).encode()).hexdigest(), 16) % 4] for i in range(length))
    with open(path, "w") as fh:
        fh.write(">%s\n%s\n" % (chrom, seq))
    return seq


def create_bam(path, chroms, lengths, reads):
    """``reads``: dicts with name, pos (0-based), seq and optional chrom_idx, cigar,
    flag, mapq, quals (list of ints), sa_tag, next_ref_id, next_pos.  Records are
    coordinate-sorted (unmapped last), as ``samtools sort`` would leave them."""
    recs = []
    for i, r in enumerate(reads):
        seq = r["seq"]
        tags = b""
        if "sa_tag" in r:
            tags += b"SAZ" + r["sa_tag"].encode() + b"\0"
        tid = r.get("chrom_idx", 0)
        enc = obam.encode_record(tid, r["pos"], r["name"], r.get("flag", 0), r.get("mapq", 60),
                                 r.get("cigar", [(0, len(seq))]), seq,
                                 r.get("quals", [40] * len(seq)), r.get("next_ref_id", -1),
                                 r.get("next_pos", -1), 0, tags)
        recs.append(((tid if tid >= 0 else 1 << 30), r["pos"], i, enc))
    recs.sort(key=lambda t: t[:3])
    obam.write_bam(path, chroms, lengths, [t[3] for t in recs])


def simple_bam(path, chrom, reads):
    """reads: (name, pos, seq[, quals[, cigar]]) tuples on one 300-bp contig."""
    out = []
    for entry in reads:
        name, pos, seq, *rest = entry
        d = {"name": name, "pos": pos, "seq": seq}
        if rest and rest[0] is not None:
            d["quals"] = rest[0]
        if len(rest) > 1:
            d["cigar"] = rest[1]
        out.append(d)
    create_bam(path, [chrom], [300], out)


def create_vcf(path, chrom, variants, sample="HG002", genotypes=None):
    """variants: (pos_1based, ref, alt[,alt2...]) tuples; alt may be a comma list."""
    with open(path, "w") as fh:
        fh.write("##fileformat=VCFv4.2\n##contig=<ID=%s,length=300>\n" % chrom)
        fh.write('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n')
        fh.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t%s\n" % sample)
        for i, (pos, ref, alt) in enumerate(variants):
            gt = genotypes[i] if genotypes else "0/1"
            fh.write("%s\t%d\t.\t%s\t%s\t.\t.\t.\tGT\t%s\n" % (chrom, pos, ref, alt, gt))
