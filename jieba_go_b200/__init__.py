"""B200-native (sm_100a) batched implementation of jieba-go's segmentation hot path.

(The directory is `jieba_go_b200` because `jieba-go_b200` is not an importable Python name.)

Product path:  tokenizer.Tokenizer -> _capi (ctypes) -> libjieba_b200.so (CUDA kernels, csrc/).
`synth` only generates inputs.  Nothing in this package imports or calls oracle/.
"""
from .tokenizer import NewJiebaTokenizer, NewTokenizer, Tokenizer  # noqa: F401
