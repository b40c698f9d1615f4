import sys, os
sys.path.insert(0,'tests'); sys.path.insert(0,'7bgzf_b200')
import helpers as H, b200bgzf as B
data = H.synth("fastq", 64 << 20)
c = B.Codec(0); m = B.MultiCodec([0, 1])
for kind, name in ((B.CONTAINER_GZIP,'gzip'),(B.CONTAINER_MIGZ,'migz'),(B.CONTAINER_GZINGA,'gzinga'),(B.CONTAINER_DICTZIP,'dictzip'),(B.CONTAINER_RAZF,'razf')):
    a = c.container(kind, data, 6); b = m.container(kind, data, 6)
    print(name, len(a), a == b, c.container_inflate(kind, b, B.VERIFY) == data)
print('bgzf', m.compress(data, 6) == c.compress(data, 6))
import subprocess
exe = os.path.join(os.path.dirname(B.APPLET_PATH), "7gzip")
r = subprocess.run([exe, "-cl6", "--devices=2"], input=data[:16<<20], capture_output=True)
print('applet --devices=2', r.returncode, r.stdout == c.container(B.CONTAINER_GZIP, data[:16<<20], 6))
