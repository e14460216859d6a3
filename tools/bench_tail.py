import sys, torch
sys.path.insert(0, "/root/repo")
from spike2former_b200 import ops
n,Q,K,h,w=16,100,150,256,256
mp=torch.randn(n,h*w,Q,device="cuda"); cls=torch.randn(n,Q,K+1,device="cuda")
for labels in (False, True):
    def run(): ops.semantic_tail(mp, cls, n=n,Q=Q,K=K,h=h,w=w,H=2*h,W=2*w, want_logits=not labels, want_labels=labels)
    for _ in range(3): run()
    torch.cuda.synchronize(); ts=[]
    for _ in range(5):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(400000); e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print("labels" if labels else "logits", sorted(ts)[2], "ms")
