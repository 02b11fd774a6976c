import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptbxl_multimodal_b200 as P
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(42)
m = P.ECGCNN(12, 256, 5).cuda().eval()
e = P.InferStep(m, B, 1000, use_graph=False)
e.xs[0].normal_(); e.xs[1].normal_()
for i in range(3): e.run(slot=i & 1)
torch.cuda.synchronize()
print('done')
