for rep in 1 2; do
for ew in 0 1; do for tn in 0 1; do
  echo -n "even_waves=$ew tune=$tn B=64: "; WFT_EVEN_WAVES=$ew WFT_TUNE=$tn ./scratch/harness 64 128 12 | tail -6 | awk '{s+=$6} END{printf "%.1f us  ", s/NR}'
  echo -n "B=256: "; WFT_EVEN_WAVES=$ew WFT_TUNE=$tn ./scratch/harness 256 128 8 | tail -4 | awk '{s+=$6} END{printf "%.1f us\n", s/NR}'
done; done; done
