for rep in 1 2; do
for v in A E B; do
  echo -n "lib$v B=64: "; LD_LIBRARY_PATH=scratch/lib$v timeout 60 ./scratch/harness 64 128 12 | tail -6 | awk '{s+=$6} END{printf "%.1f us  ", s/NR}'
  echo -n "B=256: "; LD_LIBRARY_PATH=scratch/lib$v timeout 60 ./scratch/harness 256 128 8 | tail -4 | awk '{s+=$6} END{printf "%.1f us  ", s/NR}'
  echo -n "ragged B=256: "; LD_LIBRARY_PATH=scratch/lib$v timeout 60 ./scratch/harness 256 128 8 1 | tail -4 | awk '{s+=$6} END{printf "%.1f us\n", s/NR}'
done; done
