"""Writes tests/golden/kat.json: the reference's own known-answer strings
(tests/tests.rs:22-70,100-213; doc tests src/lib.rs:14-23, src/table.rs:191-222;
examples/basic.rs) with SA/LCP values from SURVEY.md Appendix B.  The values were
transcribed from the survey (which derived them from a literal transcription of
the reference cross-checked against naive sorting); tests/test_oracle.py
re-derives every one with oracle_naive_sa and fails on any mismatch."""
import json, os
K = [
 ("apple", [0,4,3,2,1], [0,0,0,0,1]),
 ("banana", [5,3,1,0,4,2], [0,1,3,0,0,2]),
 ("mississippi", [10,7,4,1,0,9,8,6,3,5,2], [0,1,1,4,0,0,1,0,2,1,3]),
 ("tgtgtgtgcaccg", [9,8,10,11,12,7,5,3,1,6,4,2,0], [0,0,1,1,0,1,1,3,5,0,2,4,6]),
 ("", [], []),
 ("a", [0], [0]),
 ("\x00", [0], [0]),
 ("ab", [0,1], [0,0]),
 ("aa", [1,0], [0,1]),
 ("☃abc☃", [3,4,5,8,2,7,1,6,0], [0,0,0,0,1,0,2,0,3]),
 ("zzzzzaazzzzz", [5,6,11,4,10,3,9,2,8,1,7,0], [0,1,0,1,1,2,2,3,3,4,4,5]),
 ("zzzzabczzzzzabczzzzzz", [4,12,5,13,6,14,20,3,11,19,2,10,18,1,9,17,0,8,16,7,15],
  [0,8,0,7,0,6,0,1,9,1,2,10,2,3,11,3,4,12,4,5,5]),
 ("poëzie", [6,5,1,0,4,3,2], [0,0,0,0,0,0,0]),
 ("the quick brown fox was quick.",
  [9,15,3,23,19,29,21,10,7,27,2,16,1,6,26,8,28,14,12,17,4,24,11,22,0,5,25,20,13,18],
  [0,1,1,6,1,0,0,0,0,2,0,0,0,0,3,0,1,0,0,1,0,5,0,0,0,0,4,0,1,0]),
 ("The quick brown fox was very quick.",
  [9,15,3,28,23,19,34,0,21,10,7,32,2,25,16,1,6,31,8,33,14,12,17,4,29,11,26,22,5,30,24,20,13,18,27],
  [0,1,1,6,1,1,0,0,0,0,0,2,0,1,0,0,0,3,0,1,0,0,1,0,5,0,1,0,0,4,0,0,1,0,0]),
 ("aaaaaaaa", [7,6,5,4,3,2,1,0], [0,1,2,3,4,5,6,7]),
 ("abababab", [6,4,2,0,7,5,3,1], [0,2,4,6,0,1,3,5]),
 ("abcba", [4,0,3,1,2], [0,1,0,1,0]),
]
TYPES = {"banana": "LVLVLL", "mississippi": "LVLLVLLVLLL", "tgtgtgtgcaccg": "LVLVLVLLLVSSL",
         "☃abc☃": "LLLVSSLLL"}
POS = [  # (text, query, positions in SA order) tests/tests.rs:100-213 + doc tests
 ("", "", []), ("", "a", []), ("", "ab", []),
 ("a", "", []), ("a", "a", [0]), ("a", "b", []), ("a", "ab", []),
 ("ab", "", []), ("ab", "a", [0]), ("ab", "b", [1]), ("ab", "ab", [0]), ("ab", "ba", []),
 ("aa", "a", [1, 0]),
 ("zzzzzaazzzzz", "a", [5, 6]),
 ("zzzzabczzzzzabczzzzzz", "abc", [4, 12]),
 ("The quick brown fox was very quick.", "quick", [4, 29]),
 ("the quick brown fox was quick.", "quick", [4, 24]),
 ("☃abc☃", "☃", [6, 0]),
 ("The quick brown fox.", "quick", [4]),
 ("abc", "abcd", []), ("abc", "zzzz", []), ("abc", "0000", []),
]
FIX = {
 "AP009048_10000.fasta": {"n": 10001,
   "sa_sha256": "335641df720e6a760955d891723fa48fc1554248ac89a44b1a3f4a36eaa0fdc3",
   "lcp_sha256": "427e0d914a5e7c62d4b06e9b360ced03da1889f4c3fc488169e3faf83d29be57",
   "sa_head": [10000,46,9891,490,47,273,9892,7945], "lcp_head": [0,0,7,6,6,7,6,8]},
 "AP009048_100000.fasta": {"n": 100001,
   "sa_sha256": "d674074d481d76d7ac4e4ae4fe5df93a458a3b6fcb483ac92190babc52029694",
   "lcp_sha256": "10992fb21e4db240c0024acd3661b1a3af997c0fb7a1591352a89e3e1aba373d",
   "sa_head": [100000,58986,83572,20763,86246,22531,70063,46], "lcp_head": [0,0,8,12,9,9,8,7]},
}
out = {"kat": [{"text": t, "sa": s, "lcp": l} for t, s, l in K],
       "types": TYPES,
       "positions": [{"text": t, "query": q, "positions": p} for t, q, p in POS],
       "fixtures": FIX}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat.json"), "w") as f:
    json.dump(out, f, indent=1, ensure_ascii=True)
