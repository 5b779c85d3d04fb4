"""Rule-based scorers and the ``default_compute_score`` router (reference:
3rdparty/verl/verl/utils/reward_score/__init__.py, rlboost/verl_stream/utils/reward_score/__init__.py:47-54
which adds ``open-r1/OpenR1-Math-220k`` to the math family)."""
from __future__ import annotations

from . import code_exec, geo3k, gsm8k, math as math_score, math_dapo, prime_math, qa_em

MATH_SOURCES = {"lighteval/MATH", "DigitalLearningGmbH/MATH-lighteval", "HuggingFaceH4/MATH-500", "open-r1/OpenR1-Math-220k", "math"}
PRIME_MATH_SOURCES = {"numina_aops_forum", "numina_synthetic_math", "numina_amc_aime", "numina_synthetic_amc", "numina_cn_k12",
                      "numina_olympiads"}
QA_EM_SOURCES = {"searchR1_nq", "searchR1_triviaqa", "searchR1_popqa", "searchR1_hotpotqa", "searchR1_2wikimultihopqa",
                 "searchR1_musique", "searchR1_bamboogle"}


def default_compute_score(data_source: str, solution_str: str, ground_truth, extra_info=None, **kwargs) -> float:
    if data_source == "openai/gsm8k":
        return gsm8k.compute_score(solution_str, ground_truth)
    if data_source in MATH_SOURCES:
        return math_score.compute_score(solution_str, ground_truth)
    if data_source == "math_dapo" or str(data_source).startswith("aime"):
        return math_dapo.compute_score(solution_str, ground_truth)["score"]
    if data_source in PRIME_MATH_SOURCES:
        return float(prime_math.compute_score(solution_str, ground_truth)[0])
    if data_source == "hiyouga/geometry3k":
        return geo3k.compute_score(solution_str, ground_truth)
    if data_source in ("codecontests", "apps", "codeforces", "taco", "code_exec"):
        return code_exec.compute_score(solution_str, ground_truth, continuous=True)[0]
    if data_source in QA_EM_SOURCES:
        return qa_em.compute_score(solution_str, ground_truth)
    if str(data_source).startswith("synthetic/"):
        from ..synthetic import synthetic_token_reward

        ids = (extra_info or {}).get("response_ids")
        return synthetic_token_reward(ids if ids is not None else [ord(c) for c in solution_str], ground_truth, 0)
    raise NotImplementedError(f"no reward function for data_source={data_source!r}; pass custom_reward_function")
