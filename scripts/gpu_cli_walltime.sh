# wall time of the two drop-in CLIs on the GIAB mini trio (BASELINE configs 1 and 2)
G=tests/golden/giab
mkdir -p gpurun_out/cli
python -c "import torch; torch.zeros(1).cuda()"   # page in torch / create a context once
for i in 1 2; do
T0=$(date +%s.%N); python -m kmer_denovo_filter_b200.cli --child $G/HG002_child.bam --mother $G/HG004_mother.bam --father $G/HG003_father.bam --ref-fasta $G/mini_ref.fa --out-prefix gpurun_out/cli/disc --min-child-count 3 --candidate-summary tests/golden/expected_vcf/summary.txt 2> gpurun_out/cli/disc.log; echo "discovery wall $(echo "$(date +%s.%N) - $T0" | bc) s rc=$?"
T0=$(date +%s.%N); python -m kmer_denovo_filter_b200.cli --child $G/HG002_child.bam --mother $G/HG004_mother.bam --father $G/HG003_father.bam --vcf $G/candidates.vcf.gz --output gpurun_out/cli/annotated.vcf.gz --metrics gpurun_out/cli/metrics.json --summary gpurun_out/cli/summary.txt --proband-id HG002 2> gpurun_out/cli/vcf.log; echo "vcf-mode wall $(echo "$(date +%s.%N) - $T0" | bc) s"
done
grep -E "complete|finished|Anchoring|scan" gpurun_out/cli/disc.log | tail -8
diff <(cat gpurun_out/cli/disc.bed) tests/golden/expected_discovery/giab_discovery.bed && echo "BED identical to golden"
diff gpurun_out/cli/summary.txt tests/golden/expected_vcf/summary.txt && echo "VCF summary identical to golden"
ls -la gpurun_out/cli | head -20
