"""Command-line behaviour that needs no GPU: outputs this package cannot produce are refused
(exit 2), not silently skipped; compatibility flags are accepted with a warning (ADVICE round 1;
reference flags: cli.py:157-230)."""
import argparse
import logging

import pytest

from kmer_denovo_filter_b200 import cli


def _ns(**kw):
    return argparse.Namespace(**kw)


@pytest.mark.parametrize("flag", ["report", "kraken2_db", "kraken2_read_detail", "kraken2_span_bed"])
def test_unsupported_outputs_are_refused(flag, capsys):
    with pytest.raises(SystemExit) as e:
        cli._check_unsupported(_ns(**{flag: "something"}))
    assert e.value.code == 2
    err = capsys.readouterr().err
    assert "--" + flag.replace("_", "-") in err and "not part of this package" in err


def test_compatibility_flags_warn_and_pass(caplog):
    with caplog.at_level(logging.WARNING):
        cli._check_unsupported(_ns(memory=8, tmp_dir="/tmp/x", kraken2_confidence=0.1))
    text = " ".join(r.getMessage() for r in caplog.records)
    assert "--memory" in text and "--tmp-dir" in text and "kraken2" in text
    cli._check_unsupported(_ns())       # nothing set: silent


def test_discovery_parser_accepts_the_reference_flags():
    parse = getattr(cli, "parse_discovery_args", None) or getattr(cli, "parse_args", None)
    if parse is None:
        pytest.skip("no separate discovery parser")
    a = parse(["--child", "c.bam", "--mother", "m.bam", "--father", "f.bam", "--ref-fasta", "r.fa",
               "--out-prefix", "out/x", "--kmer-size", "31", "--threads", "4"])
    assert a.child == "c.bam" and a.out_prefix == "out/x" and a.kmer_size == 31
