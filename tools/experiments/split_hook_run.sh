# hook throughput with one-member calls on clusters (B200BGZF_SPLIT forces the cluster size; empty = the library's own choice)
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu.py -m gpu -x -q -k "cluster or hook or deterministic" 2>&1 | tail -3
python tools/phase_profile.py 6 128 fastq 2>&1 | head -5
./build/datagen sam 268435456 2 > /tmp/sam256.bin
export BGZF_METHOD=libdeflate6
for rep in 1 2; do for t in 1 4 8 16 32; do for s in 1 4 8 ""; do echo "rep=$rep SPLIT=$s threads=$t: $(B200BGZF_SPLIT=$s timeout 120 ./build/hook_mt 7bgzf_b200/7bgzf.so $t /tmp/sam256.bin 2 2>&1 | tail -1)"; done; done; done
