"""TEST INFRASTRUCTURE (oracle) -- literal Python restatement of jieba-go's Cut path.

Not product code: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import anything under oracle/.

This module restates /root/reference/tokenizer.go function by function with
the same observable data structures (Go maps -> dict, slices -> list), so that
the reference's own data-free golden vectors (tokenizer_test.go) can be
replayed against it verbatim.  It is deliberately slow and simple; the C
restatement in oracle/jieba_oracle.c is the fast checker and is cross-checked
against this file on randomised inputs.

Parity pinning: the Go toolchain is absent and the
reference's dict.txt / prefix_dictionary.gob / prob_emit.json are Git-LFS
stubs, so the reference itself cannot run here.  This restatement is pinned
against every data-free vector in tokenizer_test.go (TestSplitText,
TestMaxIndexProba, TestFindDagPath, TestStateTransitionRoute, TestCutHMM,
TestCutNonZh, TestBuildPrefixDict, TestAddWord); vectors that need the real
data files (TestCut, TestBuildDAG, TestCutDag, TestViterbi, TestLoadHMM) run
when JIEBA_DATA_DIR holds the files with the expected sha256
(tests/test_real_data.py) and are skipped otherwise.  Float-level parity of math.Log is unpinned by
the reference (no test holds log bits); see go_log().

Text is handled as `bytes` (a Go string is a byte sequence).  Tokens are
returned as (start, end, fffd) byte offsets into the input; `materialise`
turns them into the Go `[]string` (U+FFFD for invalid bytes, T:301-305).
"""
import math
import struct

from .unicode_tables import decode_rune, is_han, is_space

MIN_FLOAT = -3.14e100  # T:19

# T:24-29
STATE_CHANGE = {
    "B": ["E", "S"],
    "M": ["B", "M"],
    "E": ["B", "M"],
    "S": ["E", "S"],
}


# --------------------------------------------------------------------------
# math.Log (Go's portable implementation, src/math/log.go; FreeBSD e_log.c
# form).  Call sites: T:503, T:519.  Restated from SURVEY.md App. E; evaluated
# without FMA contraction (CPython never fuses).
# --------------------------------------------------------------------------
_LN2_HI = 6.93147180369123816490e-01
_LN2_LO = 1.90821492927058770002e-10
_L1 = 6.666666666666735130e-01
_L2 = 3.999999999940941908e-01
_L3 = 2.857142874366239149e-01
_L4 = 2.222219843214978396e-01
_L5 = 1.818357216161805012e-01
_L6 = 1.531383769920937332e-01
_L7 = 1.479819860511658591e-01
_SQRT2_2 = math.sqrt(2.0) / 2.0


def go_log(x: float) -> float:
    if x != x or x == math.inf:
        return x
    if x < 0:
        return math.nan
    if x == 0:
        return -math.inf
    f1, ki = math.frexp(x)
    if f1 < _SQRT2_2:
        f1 *= 2
        ki -= 1
    f = f1 - 1
    k = float(ki)
    s = f / (2 + f)
    s2 = s * s
    s4 = s2 * s2
    t1 = s2 * (_L1 + s4 * (_L3 + s4 * (_L5 + s4 * _L7)))
    t2 = s4 * (_L2 + s4 * (_L4 + s4 * _L6))
    r = t1 + t2
    hfsq = 0.5 * f * f
    return k * _LN2_HI - ((hfsq - (s * (hfsq + r) + k * _LN2_LO)) - f)


def f64_bits(x: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def encode_rune(cp: int) -> bytes:
    """Go string(rune): invalid code points become U+FFFD."""
    if cp < 0 or cp > 0x10FFFF or 0xD800 <= cp <= 0xDFFF:
        cp = 0xFFFD
    return chr(cp).encode("utf-8")


def decode_runes(b: bytes, s: int, e: int):
    """[]rune(text) / `range text`: list of (rune, byte_offset, width)."""
    out = []
    i = s
    while i < e:
        r, w = decode_rune(b, i, e)
        out.append((r, i, w))
        i += w
    return out


# --------------------------------------------------------------------------
# Regexes (T:21-22) as byte-offset run finders.
# --------------------------------------------------------------------------
def find_han_runs(b: bytes, s: int = 0, e: int = None, unicode_version: int = 15):
    """zh.FindAllIndex: maximal runs of \\p{Han} runes, as [start,end) byte offsets (T:21,154)."""
    if e is None:
        e = len(b)
    runs = []
    i = s
    cur = -1
    while i < e:
        r, w = decode_rune(b, i, e)
        if is_han(r, unicode_version):
            if cur < 0:
                cur = i
        else:
            if cur >= 0:
                runs.append([cur, i])
                cur = -1
        i += w
    if cur >= 0:
        runs.append([cur, e])
    return runs


def _is_alnum_byte(c: int) -> bool:
    return (0x30 <= c <= 0x39) or (0x41 <= c <= 0x5A) or (0x61 <= c <= 0x7A)


def find_alnum_runs(b: bytes, s: int, e: int):
    """alnum.FindAllIndex: maximal runs of [a-zA-Z0-9] bytes (T:22,290)."""
    runs = []
    i = s
    cur = -1
    while i < e:
        if _is_alnum_byte(b[i]):
            if cur < 0:
                cur = i
        else:
            if cur >= 0:
                runs.append([cur, i])
                cur = -1
        i += 1
    if cur >= 0:
        runs.append([cur, e])
    return runs


# --------------------------------------------------------------------------
# splitText (T:165-210).  Blocks are (id, start, end, doProcess) over the
# byte range [s, e) of `text`; marked indexes are absolute offsets.
# --------------------------------------------------------------------------
def split_text(s: int, e: int, marked):
    if len(marked) == 0:
        return [(0, s, e, False)]
    count = 0
    blocks = []
    prev_tail = s
    for i, pair in enumerate(marked):
        if pair[0] != prev_tail:
            blocks.append((count, prev_tail, pair[0], False))
            count += 1
        blocks.append((count, pair[0], pair[1], True))
        prev_tail = pair[1]
        count += 1
        if i == len(marked) - 1 and pair[1] != e:
            blocks.append((count, pair[1], e, False))
    return blocks


# --------------------------------------------------------------------------
# maxIndexProba (T:565-578): NOT an argmax.  items = [(index, proba), ...]
# --------------------------------------------------------------------------
def max_index_proba(items):
    prev = (-1, MIN_FLOAT)
    best = (-1, MIN_FLOAT)
    for item in items:
        if item[1] >= prev[1]:
            best = item
        prev = item
    if best[0] == -1:
        return prev
    return best


# findDagPath (T:552-562)
def find_dag_path(n_runes: int, dag_proba):
    best_path = []
    i = 0
    while 0 <= i < n_runes:
        tail = max_index_proba(dag_proba[i])
        best_path.append([i, tail[0]])
        i = tail[0]
    return best_path


class PrefixDictionary:
    """prefixDictionary (T:381-387): termFreq map[string]int + size."""

    def __init__(self):
        self.term_freq = {}  # bytes -> int
        self.size = 0

    # buildPrefixDictionary (T:340-366): prefixes added, last duplicate wins,
    # total counts every line.
    @classmethod
    def from_lines_prefix_mode(cls, lines):
        pd = cls()
        total = 0
        for line in lines:
            if isinstance(line, str):
                line = line.encode("utf-8")
            parts = line.split(b" ", 2)
            word = parts[0]
            count = _atoi(parts[1])
            total += count
            pd.term_freq[word] = count
            runes = decode_runes(word, 0, len(word))
            piece = b""
            for r, _, _ in runes[:-1]:
                piece += encode_rune(r)
                if piece not in pd.term_freq:
                    pd.term_freq[piece] = 0
        pd.size = total
        return pd

    # newPrefixDictionaryFromFile (T:389-437): no prefixes, first duplicate
    # wins and only it is counted in size.
    @classmethod
    def from_lines_file_mode(cls, lines):
        pd = cls()
        for line in lines:
            if isinstance(line, str):
                line = line.encode("utf-8")
            parts = line.split(b" ", 2)
            word = parts[0]
            count = _atoi(parts[1])
            if word not in pd.term_freq:
                pd.term_freq[word] = count
                pd.size += count
        return pd

    # addTerm (T:580-585)
    def add_term(self, term, freq: int):
        if isinstance(term, str):
            term = term.encode("utf-8")
        self.term_freq[term] = freq
        self.size += freq

    # suggestFreq (T:589-614); pieces = tk.Cut(term, false) as byte strings
    def suggest_freq(self, term: bytes, pieces):
        d_size = float(self.size)
        if d_size < 1.0:
            d_size = 1.0
        freq = 1.0
        for p in pieces:
            piece_freq = self.term_freq.get(p, 1)
            freq *= float(piece_freq) / d_size
        a = int(freq * d_size) + 1
        b = self.term_freq.get(term, 1)
        return a if a > b else b

    # buildDag (T:462-497); runes = list of code points of a Han block.
    def build_dag(self, runes):
        pieces = []
        n = len(runes)
        for i in range(n):
            key = encode_rune(runes[i])
            count = self.term_freq.get(key)
            if count is None or count == 0:
                pieces.append((i, i + 1))
                continue
            part = b""
            for j in range(n - i):
                part += encode_rune(runes[i + j])
                val = self.term_freq.get(part)
                if val is None:
                    break
                if val > 0:
                    pieces.append((i, j + 1 + i))
        dag = {}
        for p in pieces:
            dag.setdefault(p[0], []).append(p[1])
        return dag

    # calcDagProba (T:502-548)
    def calc_dag_proba(self, runes, dag):
        total = go_log(float(self.size))
        n = len(runes)
        dag_proba = {}
        for i in range(n - 1, -1, -1):
            dag_proba[i] = []
            for j in dag.get(i, []):
                tf = 1.0
                key = b"".join(encode_rune(r) for r in runes[i:j])
                val = self.term_freq.get(key)
                if val is not None:
                    tf = float(val)
                piece_freq = go_log(tf) - total
                next_piece = [(j, 0.0)]
                if j in dag_proba:
                    next_piece = dag_proba[j]
                next_best = max_index_proba(next_piece)
                piece_proba = piece_freq + next_best[1]
                dag_proba[i].append((j, piece_proba))
        return dag_proba


def _atoi(b: bytes) -> int:
    """strconv.Atoi: optional sign + decimal digits only.  A negative count is refused: Atoi takes it (T:414), but a
    rune can then have no DAG edge at all (T:468-482) and Cut walks off the DAG -- there is nothing to restate; the
    product rejects it too (JB_EFORMAT)."""
    s = b.decode("utf-8", errors="strict")
    t = s[1:] if s[:1] in "+-" else s
    if not t or not all("0" <= ch <= "9" for ch in t):
        raise ValueError("strconv.Atoi: parsing %r: invalid syntax" % s)
    v = int(s)
    if v < 0:
        raise ValueError("negative frequency %r" % s)
    return v


def split_dict_lines(data: bytes):
    """bufio.Scanner default ScanLines: split on \\n, strip one trailing \\r, drop final empty."""
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    return [ln[:-1] if ln.endswith(b"\r") else ln for ln in lines]


class HiddenMarkovModel:
    """hiddenMarkovModel (T:616-621) with newJiebaHMM's hard-coded start/trans (T:628-652)."""

    def __init__(self, emit_p, start_p=None, trans_p=None):
        self.start_p = start_p or {
            "B": -0.26268660809250016,
            "E": MIN_FLOAT,
            "M": MIN_FLOAT,
            "S": -1.4652633398537678,
        }
        self.trans_p = trans_p or {
            "B": {"E": -0.51082562376599, "M": -0.916290731874155},
            "E": {"B": -0.5897149736854513, "S": -0.8085250474669937},
            "M": {"E": -0.33344856811948514, "M": -1.2603623820268226},
            "S": {"B": -0.7211965654669841, "S": -0.6658631448798212},
        }
        # emit_p: {"B": {rune_string_or_cp: float}, ...}; keys normalised to code points.
        self.emit_p = {}
        for s in "BMES":
            tab = {}
            for k, v in emit_p.get(s, {}).items():
                if isinstance(k, str):
                    if len(k) != 1:
                        continue  # only single-rune keys are ever queried (T:689,708)
                    k = ord(k)
                tab[k] = float(v)
            self.emit_p[s] = tab
        self.route_ties = 0  # Q12: exact ties in stateTransitionRoute (Go map order) -- expect 0

    # stateTransitionRoute (T:736-756).  Go iterates a map; exact ties between
    # two routes > minFloat would be nondeterministic there (counted here and
    # broken in stateChange list order).
    def state_transition_route(self, step, now_state, hidden_states):
        routes = []
        for prev_state in STATE_CHANGE[now_state]:
            prev_prob = hidden_states[step - 1][prev_state]
            route_prob = prev_prob + self.trans_p[prev_state][now_state]
            routes.append((prev_state, route_prob))
        best_prev = ""
        best_proba = MIN_FLOAT
        for prev_state, route_proba in routes:
            if route_proba > best_proba:
                best_prev = prev_state
                best_proba = route_proba
        if routes[0][1] == routes[1][1] and routes[0][1] > MIN_FLOAT:
            self.route_ties += 1
        return best_prev, best_proba

    # viterbi (T:668-730); runes = code points of the run.
    def viterbi(self, runes):
        n = len(runes)
        if n == 1:
            return ["S"]
        hsp = {0: {}}
        full_path = {"B": ["B"], "M": ["M"], "E": ["E"], "S": ["S"]}
        states = ["B", "M", "E", "S"]
        for s in states:
            emit = self.emit_p[s].get(runes[0], MIN_FLOAT)
            hsp[0][s] = self.start_p[s] + emit
        for i in range(1, n):
            hsp[i] = {}
            partial = {}
            for s in states:
                frm, proba = self.state_transition_route(i, s, hsp)
                emit = self.emit_p[s].get(runes[i], MIN_FLOAT)
                hsp[i][s] = proba + emit
                partial[s] = list(full_path.get(frm, [])) + [s]
            full_path = partial
        e = hsp[n - 1]["E"]
        s = hsp[n - 1]["S"]
        if e > s:
            return full_path["E"]
        return full_path["S"]


# cutHMM (T:273-285): iterates the PATH, not the text -- a short path drops
# the run's tail.  Returns [(rune_start, rune_end)].
def cut_hmm(n_runes: int, path):
    pieces = []
    piece_start = 0
    for i, state in enumerate(path):
        piece_end = i + 1
        if state == "E" or state == "S":
            pieces.append((piece_start, piece_end))
            piece_start = piece_end
    return pieces


class Tokenizer:
    """Tokenizer (T:52-59) with Cut (T:151-162)."""

    def __init__(self, pd: PrefixDictionary, hmm: HiddenMarkovModel, unicode_version: int = 15):
        self.pd = pd
        self.hmm = hmm
        self.unicode_version = unicode_version

    # cutDAG (T:258-270) on rune list -> [(i, j)] rune ranges
    def cut_dag(self, runes):
        dag = self.pd.build_dag(runes)
        dag_proba = self.pd.calc_dag_proba(runes, dag)
        return find_dag_path(len(runes), dag_proba)

    # cutZh (T:221-255) -> [(rune_i, rune_j)] over the block's runes
    def cut_zh(self, runes, hmm: bool):
        dag_pieces = self.cut_dag(runes)
        if not hmm:
            return [tuple(p) for p in dag_pieces]
        words = []
        uncut = []  # rune indexes of the current run
        for idx, (a, b) in enumerate(dag_pieces):
            if b - a == 1:
                uncut.append(a)
                if idx + 1 >= len(dag_pieces) and len(uncut) != 0:
                    words.extend(self._flush_run(runes, uncut))
                    uncut = []
            else:
                if len(uncut) != 0:
                    words.extend(self._flush_run(runes, uncut))
                    uncut = []
                words.append((a, b))
        return words

    def _flush_run(self, runes, uncut):
        run = [runes[i] for i in uncut]
        v = self.hmm.viterbi(run)
        base = uncut[0]
        return [(base + a, base + b) for a, b in cut_hmm(len(run), v)]

    # cutNonZh (T:289-310) over bytes [s, e) -> [(start, end, fffd)]
    def cut_non_zh(self, b: bytes, s: int, e: int):
        alnum_idx = find_alnum_runs(b, s, e)
        if len(alnum_idx) == 0:
            return []
        out = []
        for _, bs, be, do_process in split_text(s, e, alnum_idx):
            if do_process:
                out.append((bs, be, False))
            else:
                for r, off, w in decode_runes(b, bs, be):
                    if is_space(r):
                        continue
                    # string(r): an invalid byte decodes to U+FFFD (width 1)
                    fffd = (r == 0xFFFD and w == 1)
                    out.append((off, off + w, fffd))
        return out

    # Cut (T:151-162) -> [(start, end, fffd)]
    def cut(self, text, use_hmm: bool):
        b = text.encode("utf-8") if isinstance(text, str) else bytes(text)
        zh_idx = find_han_runs(b, 0, len(b), self.unicode_version)
        blocks = split_text(0, len(b), zh_idx)
        result = []
        for _, bs, be, do_process in blocks:
            if do_process:
                rl = decode_runes(b, bs, be)
                runes = [r for r, _, _ in rl]
                offs = [o for _, o, _ in rl] + [be]
                for a, c in self.cut_zh(runes, use_hmm):
                    result.append((offs[a], offs[c], False))
            else:
                result.extend(self.cut_non_zh(b, bs, be))
        return result

    # AddWord (T:372-379) as the code intends it (the reference deadlocks: Lock at T:376, then addTerm locks at T:581)
    def add_word(self, word, freq: int):
        b = word.encode("utf-8") if isinstance(word, str) else bytes(word)
        if freq < 1:
            pieces = [b"\xef\xbf\xbd" if f else b[s:e] for s, e, f in self.cut(b, False)]
            freq = self.pd.suggest_freq(b, pieces)
        self.pd.add_term(b, freq)

    def cut_strings(self, text, use_hmm: bool):
        b = text.encode("utf-8") if isinstance(text, str) else bytes(text)
        return materialise(b, self.cut(b, use_hmm))


def materialise(b: bytes, tokens):
    """Turn (start, end, fffd) into the Go []string (as Python str)."""
    out = []
    for s, e, fffd in tokens:
        out.append("�" if fffd else b[s:e].decode("utf-8", errors="replace"))
    return out


# ---------------------------------------------------------------------------
# encoding/gob stream holding one map[string]int (prefix_dictionary.gob, T:439-458, written by
# gob.NewEncoder(f).Encode(m), tokenizer_test.go:704-712).  Wire format as published in the
# encoding/gob package documentation: a stream of length-prefixed messages; a
# message with a negative type id defines a type (skipped: key and element are the builtin
# string / int); the value message is "type id, 0 (singleton marker), count, count x (string, int)".
# ---------------------------------------------------------------------------
def _gob_uint(b: bytes, i: int):
    x = b[i]
    if x < 128:
        return x, i + 1
    n = 256 - x
    if n < 1 or n > 8 or i + 1 + n > len(b):
        raise ValueError("gob: bad unsigned integer")
    return int.from_bytes(b[i + 1:i + 1 + n], "big"), i + 1 + n


def _gob_int(b: bytes, i: int):
    u, i = _gob_uint(b, i)
    return (~(u >> 1) if u & 1 else u >> 1), i


def read_gob_map_string_int(data: bytes) -> dict:
    i, out = 0, None
    while i < len(data):
        mlen, i = _gob_uint(data, i)
        end = i + mlen
        if end > len(data):
            raise ValueError("gob: truncated message")
        tid, j = _gob_int(data, i)
        if tid >= 0:
            if out is not None:
                raise ValueError("gob: more than one value")
            marker, j = _gob_uint(data, j)
            if marker != 0:
                raise ValueError("gob: expected the singleton marker")
            count, j = _gob_uint(data, j)
            out = {}
            for _ in range(count):
                kl, j = _gob_uint(data, j)
                key = data[j:j + kl]
                j += kl
                val, j = _gob_int(data, j)
                out[key] = val
            if j != end:
                raise ValueError("gob: malformed map payload")
        i = end
    if out is None:
        raise ValueError("gob: no value")
    return out
