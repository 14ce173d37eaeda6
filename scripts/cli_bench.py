"""End-to-end wall clock of the rdp_classifier executable (FASTA file in, text file out) on the bench workload."""
import sys, time, subprocess, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'pangea-plus_b200'))
import numpy as np
from pangea_b200 import synth
tr=synth.synth16s(0x9178, 9178, 1219)
names, anc = tr["node_names"], tr["anc"]
os.makedirs('/tmp/cli', exist_ok=True)
with open('/tmp/cli/train.fa','w') as f:
    for i,g in enumerate(tr["genus"]):
        s=tr["data"][tr["off"][i]:tr["off"][i+1]].tobytes().decode()
        f.write(f">T{i:06d}\t"+";".join(names[n] for n in anc[g])+"\n"+s+"\n")
n=262144
data,off,src=synth.synth_reads(0x250,tr,n,paired=True)
L=689
arr=data.reshape(n,L)
t=time.time()
with open('/tmp/cli/q.fa','wb') as f:
    for i in range(n):
        f.write(b">r%07d:AB\n"%i); f.write(arr[i].tobytes()); f.write(b"\n")
print("wrote query", time.time()-t)
B=os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'pangea-plus_b200', 'bin', 'rdp_classifier')
t=time.time(); subprocess.run([B,'--train','/tmp/cli/train.fa','-t','/tmp/cli/m.pgm'],check=True); print("train", time.time()-t)
for fmt in ('allrank','pangea'):
    t=time.time(); subprocess.run([B,'-q','/tmp/cli/q.fa','-o','/tmp/cli/o_%s.txt'%fmt,'-t','/tmp/cli/m.pgm','-f',fmt],check=True,env=dict(os.environ,PG_TIMING='1')); dt=time.time()-t
    print(fmt, "classify CLI wall %.2f s -> %.0f reads/s"%(dt, n/dt), os.path.getsize('/tmp/cli/o_%s.txt'%fmt))
