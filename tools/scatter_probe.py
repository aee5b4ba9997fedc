import sys, torch
sys.path.insert(0, "/root/repo")
import __graft_entry__ as e; e.build()
import audio_visual_llm_b200 as pkg
L = pkg._lib
N = int(sys.argv[1]); dt = torch.float32 if sys.argv[2] == "f32" else torch.bfloat16
B, P, K, H = 2, 2, 64, 128
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
a = torch.randn(B * N, K, generator=g).bfloat16().to(dev)
w = (torch.randn(H, K, generator=g) / 8).bfloat16().to(dev)
emb = torch.zeros(B, P + N, H, dtype=dt, device=dev)
L.proj_fwd([a], [w], emb[:, P:, :])
torch.cuda.synchronize()
ref = (a.double() @ w.double().T).view(B, N, H)
print(N, sys.argv[2], "max err", float((emb[:, P:].double() - ref).abs().max()), "prompt rows untouched", float(emb[:, :P].abs().max()) == 0.0)
