# 1 GPU: the placement kernel in its exchange form (one virtual rank, blocks in place / through the buffers) against the
# same reads on the whole DB -- why is a read with ambiguity codes dearer in the exchange form?
mkdir -p gpurun_out
for amb in "" "--no-ambiguity"; do
  echo "== $amb in place"; RP_XCHG_DEBUG=1 RP_DEBUG_GEOM=1 timeout 300 python tools/xchg_local_bench.py --k 13 --world 1 --reads 100000 --whole $amb 2>&1 | grep -v "^rp_xchg\[0\] \(h2d\|sub-batch plan\|fill\|answers\)" | tail -14
done
echo "== amb through the buffers"; RP_XCHG_COPY_LOCAL=1 RP_XCHG_DEBUG=1 timeout 300 python tools/xchg_local_bench.py --k 13 --world 1 --reads 100000 2>&1 | grep "pipeline\|sub-batch [0-9]" | tail -8
echo "== amb, unbatched build forced"; RP_AMB_BATCH=0 RP_XCHG_DEBUG=1 timeout 300 python tools/xchg_local_bench.py --k 13 --world 1 --reads 100000 --whole 2>&1 | grep "pipeline\|whole" | tail -3
