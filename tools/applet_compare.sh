#!/bin/bash
# "As shipped" comparison: the reference applet vs ours, stdin/stdout on tmpfs, 1 GiB FASTQ-like.
set -e
cd "$(dirname "$0")/.."
D=/dev/shm/b200bgzf_cmp; mkdir -p $D
MIB=${1:-1024}
./build/datagen fastq $((MIB*1048576)) 1 | head -c $((MIB*1048576)) > $D/in.fq
N=$(nproc)
echo "cores=$N input=$(stat -c %s $D/in.fq)"
t() { local s=$(date +%s.%N); "$@"; local e=$(date +%s.%N); echo "$(echo "$e - $s" | bc -l 2>/dev/null || python3 -c "print($e-$s)")"; }
for th in 1 $N; do
  s=$(date +%s.%N); ./oracle/_ref/7bgzf -c -l6 -@$th < $D/in.fq > $D/ref.bgz 2>/dev/null; e=$(date +%s.%N)
  python3 -c "print('reference 7bgzf -c -l6 -@$th : %.2f s  %.3f GB/s  size %d' % ($e-$s, $MIB*1048576/($e-$s)/1e9, $(stat -c %s $D/ref.bgz)))"
done
s=$(date +%s.%N); ./7bgzf_b200/7bgzf -c -l6 < $D/in.fq > $D/gpu.bgz 2>$D/gpu.err; e=$(date +%s.%N)
python3 -c "print('b200 7bgzf -c -l6 (incl. CUDA init) : %.2f s  %.3f GB/s  size %d' % ($e-$s, $MIB*1048576/($e-$s)/1e9, $(stat -c %s $D/gpu.bgz)))"
tail -2 $D/gpu.err
for th in 1 $N; do
  s=$(date +%s.%N); ./oracle/_ref/7bgzf -d -@$th < $D/ref.bgz > $D/out.fq 2>/dev/null; e=$(date +%s.%N)
  python3 -c "print('reference 7bgzf -d -@$th : %.2f s  %.3f GB/s' % ($e-$s, $MIB*1048576/($e-$s)/1e9))"
done
s=$(date +%s.%N); ./7bgzf_b200/7bgzf -d < $D/ref.bgz > $D/out2.fq 2>/dev/null; e=$(date +%s.%N)
python3 -c "print('b200 7bgzf -d (reference stream) : %.2f s  %.3f GB/s' % ($e-$s, $MIB*1048576/($e-$s)/1e9))"
cmp $D/out2.fq $D/in.fq && echo "b200 -d of reference stream: identical to input"
./oracle/_ref/7bgzf -d -@$N < $D/gpu.bgz 2>/dev/null | cmp - $D/in.fq && echo "reference -d of b200 stream: identical to input"
rm -rf $D
