#!/bin/bash
# A/B measurement tooling (not part of the product): builds a second copy of libfusionsim.so under
# tools/scratch/ab/<name>/ from a git revision (or the working tree, rev = WORK) with extra nvcc flags.
#   tools/ab_build.sh <name> <rev|WORK> [EXTRA flags]      then FSIM_LIB_PATH=tools/scratch/ab/<name>/fusion_sim_b200/csrc/libfusionsim.so
set -e
cd "$(dirname "$0")/.."
name=$1; rev=$2; shift 2
d=tools/scratch/ab/$name
rm -rf "$d"; mkdir -p "$d"
if [ "$rev" = WORK ]; then
  mkdir -p "$d/fusion_sim_b200"; cp -r include "$d/"; cp -r fusion_sim_b200/csrc "$d/fusion_sim_b200/"
  rm -f "$d"/fusion_sim_b200/csrc/*.o "$d"/fusion_sim_b200/csrc/*.so
else
  git archive "$rev" fusion_sim_b200/csrc include | tar -x -C "$d"
fi
make -C "$d/fusion_sim_b200/csrc" -j8 libfusionsim.so EXTRA="$*" >/dev/null
rm -f "$d"/fusion_sim_b200/csrc/*.o
echo "$d/fusion_sim_b200/csrc/libfusionsim.so"
