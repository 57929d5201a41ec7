"""TEST INFRASTRUCTURE (oracle) -- Unicode facts the jieba-go Cut path depends on.

Not product code: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import anything under oracle/.

The reference gets these from the Go standard library, which is not vendored
under /root/reference:
  * regexp `\\p{Han}`            -- /root/reference/tokenizer.go:21   (zh)
  * unicode.IsSpace             -- /root/reference/tokenizer.go:302
  * UTF-8 decoding of `range s` -- /root/reference/tokenizer.go:301
The Han table depends on the Go toolchain's Unicode version (go.mod pins only
`go 1.18`): Go 1.18-1.20 ship Unicode 13.0, Go >= 1.21 ships Unicode 15.0.
Both are restated here; 15.0 is the default.  Only common CJK is pinned by the
reference's own tests (tokenizer_test.go:66-71); the fringe is unpinned.
"""

# Script=Han, Unicode 13.0 (Go 1.18 - 1.20).
HAN_RANGES_13 = [
    (0x2E80, 0x2E99), (0x2E9B, 0x2EF3), (0x2F00, 0x2FD5), (0x3005, 0x3005),
    (0x3007, 0x3007), (0x3021, 0x3029), (0x3038, 0x303B), (0x3400, 0x4DBF),
    (0x4E00, 0x9FFC), (0xF900, 0xFA6D), (0xFA70, 0xFAD9), (0x16FE3, 0x16FE3),
    (0x16FF0, 0x16FF1), (0x20000, 0x2A6DD), (0x2A700, 0x2B734),
    (0x2B740, 0x2B81D), (0x2B820, 0x2CEA1), (0x2CEB0, 0x2EBE0),
    (0x2F800, 0x2FA1D), (0x30000, 0x3134A),
]

# Script=Han, Unicode 15.0 (Go >= 1.21).
HAN_RANGES_15 = [
    (0x2E80, 0x2E99), (0x2E9B, 0x2EF3), (0x2F00, 0x2FD5), (0x3005, 0x3005),
    (0x3007, 0x3007), (0x3021, 0x3029), (0x3038, 0x303B), (0x3400, 0x4DBF),
    (0x4E00, 0x9FFF), (0xF900, 0xFA6D), (0xFA70, 0xFAD9), (0x16FE2, 0x16FE3),
    (0x16FF0, 0x16FF1), (0x20000, 0x2A6DF), (0x2A700, 0x2B739),
    (0x2B740, 0x2B81D), (0x2B820, 0x2CEA1), (0x2CEB0, 0x2EBE0),
    (0x2F800, 0x2FA1D), (0x30000, 0x3134A), (0x31350, 0x323AF),
]

HAN_RANGES = {13: HAN_RANGES_13, 15: HAN_RANGES_15}

# unicode.IsSpace (White_Space property; stable across Unicode versions).
SPACE_CODEPOINTS = frozenset(
    [0x09, 0x0A, 0x0B, 0x0C, 0x0D, 0x20, 0x85, 0xA0, 0x1680]
    + list(range(0x2000, 0x200B))
    + [0x2028, 0x2029, 0x202F, 0x205F, 0x3000]
)


def is_han(cp: int, version: int = 15) -> bool:
    for lo, hi in HAN_RANGES[version]:
        if lo <= cp <= hi:
            return True
    return False


def is_space(cp: int) -> bool:
    return cp in SPACE_CODEPOINTS


def decode_rune(b: bytes, i: int, end: int):
    """Go's utf8.DecodeRune on b[i:end]: returns (rune, width).

    Anything that is not a well-formed sequence yields (U+FFFD, 1).
    """
    n = end - i
    if n < 1:
        return 0xFFFD, 0
    b0 = b[i]
    if b0 < 0x80:
        return b0, 1
    if 0xC2 <= b0 <= 0xDF:
        if n >= 2 and 0x80 <= b[i + 1] <= 0xBF:
            return ((b0 & 0x1F) << 6) | (b[i + 1] & 0x3F), 2
        return 0xFFFD, 1
    if 0xE0 <= b0 <= 0xEF:
        lo, hi = 0x80, 0xBF
        if b0 == 0xE0:
            lo = 0xA0
        elif b0 == 0xED:
            hi = 0x9F
        if n >= 3 and lo <= b[i + 1] <= hi and 0x80 <= b[i + 2] <= 0xBF:
            return ((b0 & 0x0F) << 12) | ((b[i + 1] & 0x3F) << 6) | (b[i + 2] & 0x3F), 3
        return 0xFFFD, 1
    if 0xF0 <= b0 <= 0xF4:
        lo, hi = 0x80, 0xBF
        if b0 == 0xF0:
            lo = 0x90
        elif b0 == 0xF4:
            hi = 0x8F
        if (n >= 4 and lo <= b[i + 1] <= hi and 0x80 <= b[i + 2] <= 0xBF
                and 0x80 <= b[i + 3] <= 0xBF):
            return (((b0 & 0x07) << 18) | ((b[i + 1] & 0x3F) << 12)
                    | ((b[i + 2] & 0x3F) << 6) | (b[i + 3] & 0x3F)), 4
        return 0xFFFD, 1
    return 0xFFFD, 1
