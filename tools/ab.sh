#!/bin/bash
# usage: ab.sh out label=ENV... ; each arg "label|ENV1=a ENV2=b"
out=$1; shift
for rep in 1 2; do
for e in "$@"; do
  label=${e%%|*}; envs=${e#*|}
  echo "# $label" >> $out
  env $envs timeout 300 python tools/prof_run.py 3000 3 >> $out 2>&1
done
done
