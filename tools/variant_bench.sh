#!/bin/bash
# run the sketch bench for every kernel-variant library under gpurun_variants/ (experiments)
for f in gpurun_variants/lib_*.so; do
  echo -n "$f  "
  HG_LIB=$PWD/$f python bench.py --steps 6 --warmup 3 --no-dist --no-cpu-baseline --genomes ${GENOMES:-400} 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['stages_ms']['kmer_hash'])"
done
