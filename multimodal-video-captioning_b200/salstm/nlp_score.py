"""Caption metrics behind ``losses.NLPScore`` (reference src/losses.py:140-160, which wraps the COCO caption
evaluation scorers of pycocoevalcap on dicts ``{video_id: [sentence, ...]}``): corpus BLEU-1..4 (clipped n-gram
precision, closest-reference brevity penalty), ROUGE-L (LCS F-measure, beta = 1.2, best precision / recall over the
references) and CIDEr (tf-idf weighted n-gram cosine, n = 1..4, Gaussian length penalty sigma = 6, x10), restated
from the published definitions (Papineni et al. 2002; Lin 2004; Vedantam et al. 2015) and pinned against the
reference's own scorers on seeded sentence sets (tools/make_golden_r2.py -> tests/golden/nlp_scores_small.json).

Why here: `Trainer.eval` runs every epoch (train.py:316-348, eval_freq = 1) and, once decoding takes milliseconds, the
scorer is the epoch-time floor.  The implementations below tokenise by whitespace exactly like the
reference's call path (NLPScore hands raw strings to the scorers; no PTB tokeniser) and use Counter arithmetic over
integer-hashed n-grams.  METEOR needs the Java jar shipped with pycocoevalcap: it is delegated to the reference's
scorer when both are present and reported as NaN otherwise (Trainer.eval logs BLEU / ROUGE_L / CIDEr only).
"""
from __future__ import annotations

import math
import shutil
from collections import Counter
from typing import Dict, List, Sequence, Tuple


def _ngrams(tokens: Sequence[str], n_max: int = 4) -> Counter:
    c: Counter = Counter()
    L = len(tokens)
    for n in range(1, n_max + 1):
        for i in range(L - n + 1):
            c[tuple(tokens[i:i + n])] += 1
    return c


def _check(gts: Dict, res: Dict) -> List:
    ids = sorted(gts.keys())
    if ids != sorted(res.keys()):
        raise AssertionError("reference and hypothesis dictionaries must have the same keys")
    for i in ids:
        if not (isinstance(res[i], list) and len(res[i]) == 1 and isinstance(gts[i], list) and len(gts[i]) >= 1):
            raise AssertionError("every id needs exactly one hypothesis and at least one reference sentence")
    return ids


def bleu(gts: Dict, res: Dict, n: int = 4) -> Tuple[List[float], List[List[float]]]:
    """Corpus-level BLEU-1..n with the `closest` effective reference length -> (scores, per-sentence scores)."""
    ids = _check(gts, res)
    tiny, small = 1e-15, 1e-9
    tot_guess, tot_correct = [0] * n, [0] * n
    tot_test = tot_ref = 0
    per_sentence: List[List[float]] = [[] for _ in range(n)]
    for i in ids:
        hyp = res[i][0].split()
        refs = [r.split() for r in gts[i]]
        counts = _ngrams(hyp, n)
        max_ref: Counter = Counter()
        for r in refs:
            for g, c in _ngrams(r, n).items():
                if c > max_ref[g]:
                    max_ref[g] = c
        testlen = len(hyp)
        reflen = min((abs(len(r) - testlen), len(r)) for r in refs)[1]        # closest; ties -> the shorter one
        guess = [max(0, testlen - k) for k in range(n)]
        correct = [0] * n
        for g, c in counts.items():
            correct[len(g) - 1] += min(c, max_ref.get(g, 0))
        tot_test += testlen
        tot_ref += reflen
        b = 1.0
        for k in range(n):
            tot_guess[k] += guess[k]
            tot_correct[k] += correct[k]
            b *= (correct[k] + tiny) / (guess[k] + small)
            per_sentence[k].append(b ** (1.0 / (k + 1)))
        ratio = (testlen + tiny) / (reflen + small)
        if ratio < 1:
            for k in range(n):
                per_sentence[k][-1] *= math.exp(1 - 1 / ratio)
    scores = []
    b = 1.0
    for k in range(n):
        b *= (tot_correct[k] + tiny) / (tot_guess[k] + small)
        scores.append(b ** (1.0 / (k + 1)))
    ratio = (tot_test + tiny) / (tot_ref + small)
    if ratio < 1:
        scores = [s * math.exp(1 - 1 / ratio) for s in scores]
    return scores, per_sentence


def _lcs(a: Sequence[str], b: Sequence[str]) -> int:
    if len(a) < len(b):
        a, b = b, a
    prev = [0] * (len(b) + 1)
    for x in a:
        cur = [0]
        for j, y in enumerate(b, 1):
            cur.append(prev[j - 1] + 1 if x == y else max(prev[j], cur[j - 1]))
        prev = cur
    return prev[-1]


def rouge_l(gts: Dict, res: Dict, beta: float = 1.2) -> Tuple[float, List[float]]:
    ids = _check(gts, res)
    out = []
    for i in ids:
        hyp = res[i][0].split(" ")
        prec, rec = [], []
        for r in gts[i]:
            ref = r.split(" ")
            lcs = _lcs(ref, hyp)
            prec.append(lcs / float(len(hyp)))
            rec.append(lcs / float(len(ref)))
        p, r_ = max(prec), max(rec)
        out.append(((1 + beta ** 2) * p * r_) / float(r_ + beta ** 2 * p) if p != 0 and r_ != 0 else 0.0)
    return sum(out) / len(out), out


def cider(gts: Dict, res: Dict, n: int = 4, sigma: float = 6.0) -> Tuple[float, List[float]]:
    ids = _check(gts, res)
    hyps = [_ngrams(res[i][0].split(), n) for i in ids]
    refs = [[_ngrams(r.split(), n) for r in gts[i]] for i in ids]
    df: Counter = Counter()
    for rs in refs:
        seen = set()
        for r in rs:
            seen.update(r.keys())
        for g in seen:
            df[g] += 1
    log_n = math.log(float(len(ids)))

    def vec(counts: Counter):
        v = [dict() for _ in range(n)]
        norm = [0.0] * n
        length = 0
        for g, tf in counts.items():
            k = len(g) - 1
            w = float(tf) * (log_n - math.log(max(1.0, df.get(g, 0))))
            v[k][g] = w
            norm[k] += w * w
            if k == 1:
                length += tf
        return v, [math.sqrt(x) for x in norm], length

    def sim(vh, vr, nh, nr, lh, lr):
        delta = float(lh - lr)
        val = [0.0] * n
        for k in range(n):
            for g, w in vh[k].items():
                val[k] += min(w, vr[k].get(g, 0.0)) * vr[k].get(g, 0.0)
            if nh[k] != 0 and nr[k] != 0:
                val[k] /= nh[k] * nr[k]
            val[k] *= math.e ** (-(delta ** 2) / (2 * sigma ** 2))
        return val

    out = []
    for h, rs in zip(hyps, refs):
        vh, nh, lh = vec(h)
        score = [0.0] * n
        for r in rs:
            vr, nr, lr = vec(r)
            s = sim(vh, vr, nh, nr, lh, lr)
            score = [a + b for a, b in zip(score, s)]
        out.append(sum(score) / n / len(rs) * 10.0)
    return sum(out) / len(out), out


def meteor(gts: Dict, res: Dict, reference_losses=None) -> float:
    """METEOR 1.5 is a Java program driven through a pipe (pycocoevalcap/meteor): delegated to the reference's scorer
    when it and a `java` binary are available, NaN otherwise."""
    if reference_losses is None or shutil.which("java") is None:
        return float("nan")
    try:
        scorer = reference_losses.Meteor()
        score, _ = scorer.compute_score(gts, res)
        return float(score)
    except Exception:
        return float("nan")


def nlp_score(ref: Dict, hypo: Dict, reference_losses=None) -> Dict[str, float]:
    """Same dictionary as the reference's NLPScore(ref, hypo): Bleu_1..4, METEOR, ROUGE_L, CIDEr."""
    b, _ = bleu(ref, hypo, 4)
    out = {f"Bleu_{k + 1}": b[k] for k in range(4)}
    out["METEOR"] = meteor(ref, hypo, reference_losses)
    out["ROUGE_L"] = rouge_l(ref, hypo)[0]
    out["CIDEr"] = cider(ref, hypo)[0]
    return out
