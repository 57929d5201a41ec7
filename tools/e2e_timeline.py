import time, numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jieba_go_b200 import synth
from jieba_go_b200.tokenizer import Tokenizer
sd = synth.make_dictionary(n_words=349000, seed=synth.SEED_BASE)
emit = synth.make_emit(sd)
text, doc_off = synth.make_corpus(sd, 'freq', 1_000_000_000, synth.SEED_BASE + 2, device='cuda')
h_text = text.cpu().pin_memory(); h_np = h_text.numpy(); h_off = doc_off.cpu().numpy().astype(np.uint64)
tk = Tokenizer.from_dict_text(sd.dict_txt(), 1, emit, device=0)
for _ in range(2): tk.cut_batch_view(h_np, h_off, False).close()
os.environ["JB_TIMELINE"] = "1"
torch.cuda.synchronize(); t0 = time.perf_counter()
with tk.cut_batch_view(h_np, h_off, False) as r: n = r.n_tokens
torch.cuda.synchronize(); print('total ms', (time.perf_counter() - t0) * 1e3, n)
