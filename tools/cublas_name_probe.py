import torch
dev = torch.device("cuda:0")
M, K, H = 12000, 6144, 4096
A = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(H, K, device=dev) / 78).bfloat16()
Y = torch.empty(M, H, dtype=torch.bfloat16, device=dev); dY = torch.randn(M, H, device=dev).bfloat16()
dWb = torch.empty(H, K, dtype=torch.bfloat16, device=dev)
for _ in range(3):
    torch.matmul(A, W.t(), out=Y); torch.matmul(dY.t(), A, out=dWb)
torch.cuda.synchronize()
