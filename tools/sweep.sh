#!/bin/bash
# usage: tools/sweep.sh out.jsonl "ENV1=a ENV2=b" "ENV1=c" ...   (one quick bench run per environment string)
out=$1; shift
for e in "$@"; do
  echo "# $e" >> $out
  env $e timeout 300 python bench.py --quick --steps 2 --warmup 1 >> $out 2>> $out.err || echo "{\"failed\": \"$e\"}" >> $out
done
