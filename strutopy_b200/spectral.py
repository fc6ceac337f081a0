"""Spectral initialisation of beta (stm.py:30-296) — SURVEY.md §8f-1, a "next" row: not built yet."""


def spectral_init(corpus, K, V, maxV=5000):
    raise NotImplementedError(
        "init_type='spectral' (stm.py:30-296) is a 'next' row of the hot-path scope table (SURVEY.md "
        "§8f-1) and is not implemented yet; use init_type='random' or assign model.beta")
