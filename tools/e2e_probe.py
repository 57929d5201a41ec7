import time, numpy as np, torch, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jieba_go_b200 import synth
from jieba_go_b200.tokenizer import Tokenizer
sd = synth.make_dictionary(n_words=349000, seed=synth.SEED_BASE)
emit = synth.make_emit(sd)
text, doc_off = synth.make_corpus(sd, 'freq', 1_000_000_000, synth.SEED_BASE + 2, device='cuda')
h_text = text.cpu().pin_memory(); h_np = h_text.numpy(); h_off = doc_off.cpu().numpy().astype(np.uint64)
for mb in (64 << 20, 128 << 20, 256 << 20):
    tk = Tokenizer.from_dict_text(sd.dict_txt(), 1, emit, device=0, max_batch_bytes=mb)
    tk.cut_batch_view(h_np, h_off, False).close()
    tk.cut_batch_view(h_np, h_off, False).close()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        with tk.cut_batch_view(h_np, h_off, False) as r: n = r.n_tokens
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print('max_batch', mb, 'ms', dt * 1e3, 'GB/s', h_np.size / dt / 1e9, n)
    tk.close()
