"""Timeline of one end-to-end call (JB_TIMELINE=1 makes the library print, per sub-batch, when its H2D copy, kernels
and D2H copies completed):  python tools/e2e_timeline.py [bits|arrays] [freq|oov|long] [max_batch_bytes]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jieba_go_b200 import synth
from jieba_go_b200.tokenizer import Tokenizer
fmt = sys.argv[1] if len(sys.argv) > 1 else "bits"
kind = sys.argv[2] if len(sys.argv) > 2 else "freq"
sd = synth.make_dictionary(n_words=349000, seed=synth.SEED_BASE)
emit = synth.make_emit(sd)
text, doc_off = synth.make_corpus(sd, kind, 1_000_000_000, synth.SEED_BASE + 2, device='cuda')
h_text = text.cpu().pin_memory(); h_np = h_text.numpy(); h_off = doc_off.cpu().numpy().astype(np.uint64)
mb = int(sys.argv[3]) if len(sys.argv) > 3 else 0
tk = Tokenizer.from_dict_text(sd.dict_txt(), 1, emit, device=0, **({"max_batch_bytes": mb} if mb else {}))
call = (lambda: tk.cut_batch_bits(h_np, h_off, kind != "freq")) if fmt == "bits" else (lambda: tk.cut_batch_view(h_np, h_off, kind != "freq"))
for _ in range(2): call().close()
os.environ["JB_TIMELINE"] = "1"
torch.cuda.synchronize(); t0 = time.perf_counter()
with call() as r: n = r.n_tokens
torch.cuda.synchronize(); print('total ms', (time.perf_counter() - t0) * 1e3, n)
