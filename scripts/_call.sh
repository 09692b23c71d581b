for v in "X=0" "IPSR_BWD_TILE_KB=16 IPSR_BWD_STAGES=2" "IPSR_BWD_TILE_KB=16 IPSR_BWD_STAGES=4" "IPSR_BWD_TILE_KB=8 IPSR_BWD_STAGES=4" "IPSR_BWD_LIGHT=32" "IPSR_BWD_LIGHT=48"; do
echo "$v"; env $v python scripts/bwd_timing.py 16 256 32 2>&1 | tail -1
done
