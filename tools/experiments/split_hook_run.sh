# hook throughput, oversubscribed callers: does the number of hardware work queues matter?
O=gpurun_out; mkdir -p $O
./build/datagen sam 268435456 2 > /tmp/sam256.bin
export BGZF_METHOD=libdeflate6
for rep in 1 2 3 4 5; do for t in 32 64 128; do for c in 8 32; do echo "rep=$rep CONN=$c threads=$t: $(CUDA_DEVICE_MAX_CONNECTIONS=$c timeout 120 ./build/hook_mt 7bgzf_b200/7bgzf.so $t /tmp/sam256.bin 2 2>&1 | tail -1)"; done; done; done
