# diagnostics: row hash of column batch k on a second stream beside the LDE of batch k+1
mkdir -p gpurun_out
python bench.py --quick --steps 5 --no-cpu-baseline 2>/dev/null | cut -c1-58
for v in "--overlap-hash 1 --hash-blocks-per-sm 1" "--overlap-hash 1 --hash-blocks-per-sm 2" "--overlap-hash 1 --hash-blocks-per-sm 4" "--overlap-hash 1 --hash-blocks-per-sm 2 --lde-batch-mb 512"; do
  echo "$v"; python bench.py --quick --steps 5 --no-cpu-baseline $v 2>/dev/null | cut -c1-58
done
