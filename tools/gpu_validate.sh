# end-to-end validation against the reference's published series (development aid)
set -x
mkdir -p gpurun_out
for c in "$@"; do
  name=$(echo $c | tr '/' '_')
  timeout 1500 python tools/validate_published.py $c > gpurun_out/validate_$name.log 2>&1; echo "rc=$?" >> gpurun_out/validate_$name.log
done
tail -n 15 gpurun_out/validate_*.log
