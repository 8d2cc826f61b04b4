"""The retrieval scenarios recorded in the golden file — one definition used by make_golden.py (real
reference code) and by the tests (this repo's code)."""
from .corpus import COMPANY, OTHER, QUESTIONS

MQ_ALL = {"synonym": True, "subquestion": True, "variant": True}

# name -> kwargs of VectorRetriever.retrieve_by_company_name (constructor kwargs under "_init")
VECTOR_CASES = {
    "basic_plain": dict(company_name=COMPANY, query=QUESTIONS[0], top_n=30, use_hyde=False, use_multi_query=False),
    "basic_years_2024": dict(company_name=COMPANY, query=QUESTIONS[0], top_n=10, use_hyde=False, use_multi_query=False,
                             selected_years=[2024]),
    "basic_years_3": dict(company_name=COMPANY, query=QUESTIONS[1], top_n=10, use_hyde=False, use_multi_query=False,
                          selected_years=[2023, 2024, 2025]),
    "basic_years_missing_falls_back": dict(company_name=COMPANY, query=QUESTIONS[1], top_n=5, use_hyde=False,
                                           use_multi_query=False, selected_years=[1999]),
    "basic_year_field_string": dict(company_name=COMPANY, query=QUESTIONS[2], top_n=5, use_hyde=False,
                                    use_multi_query=False, selected_years=[2022]),
    "basic_expanded_pages": dict(company_name=COMPANY, query=QUESTIONS[0], top_n=10, return_parent_pages=True,
                                 use_hyde=True, use_multi_query=True, multi_query_config=MQ_ALL),
    "basic_expanded_chunks": dict(company_name=COMPANY, query=QUESTIONS[3], top_n=12, use_hyde=True,
                                  use_multi_query=True, multi_query_config={"synonym": True, "subquestion": True}),
    "multi_query_no_methods": dict(company_name=COMPANY, query=QUESTIONS[2], top_n=4, use_hyde=False,
                                   use_multi_query=True, multi_query_config={"synonym": False}),
    "other_company_subset_year": dict(company_name=OTHER, query=QUESTIONS[2], top_n=50, use_hyde=False,
                                      use_multi_query=False, selected_years=[2021]),
    "ssg": dict(company_name=COMPANY, query=QUESTIONS[1], top_n=10, use_hyde=False, use_multi_query=False,
                retrieval_method="ssg", max_hops=3, neighbor_k=8),
    "triangulation": dict(company_name=COMPANY, query=QUESTIONS[1], top_n=10, use_hyde=False, use_multi_query=False,
                          retrieval_method="triangulation", max_hops=3, neighbor_k=8, selected_years=[2023, 2024]),
    "triangulation_pages_expanded": dict(company_name=COMPANY, query=QUESTIONS[3], top_n=8, return_parent_pages=True,
                                         use_hyde=True, use_multi_query=False, retrieval_method="triangulation",
                                         max_hops=2, neighbor_k=5, selected_years=[2025]),
    "hybrid_expansion": dict(company_name=COMPANY, query=QUESTIONS[0], top_n=15, use_hyde=False, use_multi_query=False,
                             retrieval_method="hybrid_expansion", max_hops=2, neighbor_k=6, selected_years=[2024]),
    "unknown_method_falls_back": dict(company_name=COMPANY, query=QUESTIONS[2], top_n=5, use_hyde=False,
                                      use_multi_query=False, retrieval_method="nonsense", selected_years=[2020]),
    # graph expansions over several reports and several (expanded) queries: many walks advance together
    "hybrid_expansion_multi": dict(company_name=COMPANY, query=QUESTIONS[3], top_n=20, return_parent_pages=True,
                                   use_hyde=True, use_multi_query=True, multi_query_config=MQ_ALL,
                                   retrieval_method="hybrid_expansion", max_hops=3, neighbor_k=6,
                                   selected_years=[2023, 2024, 2025]),
    "ssg_all_reports_expanded": dict(company_name=COMPANY, query=QUESTIONS[0], top_n=15, use_hyde=False,
                                     use_multi_query=True, multi_query_config=MQ_ALL, retrieval_method="ssg",
                                     max_hops=4, neighbor_k=10),
    "triangulation_all_reports_hyde": dict(company_name=COMPANY, query=QUESTIONS[2], top_n=12, use_hyde=True,
                                           use_multi_query=False, retrieval_method="triangulation", max_hops=3,
                                           neighbor_k=7),
    "instance_attr_overrides_default": dict(company_name=COMPANY, query=QUESTIONS[1], top_n=6, use_hyde=False,
                                            use_multi_query=False, selected_years=[2021],
                                            _init=dict(retrieval_method="triangulation", max_hops=1, neighbor_k=4)),
}

HYBRID_CASES = {
    "hybrid_default": dict(company_name=COMPANY, query=QUESTIONS[0], llm_reranking_sample_size=28, top_n=6,
                           use_hyde=False, use_multi_query=False),
    "hybrid_expanded_pages": dict(company_name=COMPANY, query=QUESTIONS[2], llm_reranking_sample_size=20, top_n=10,
                                  return_parent_pages=True, use_hyde=True, use_multi_query=True,
                                  multi_query_config=MQ_ALL, selected_years=[2023, 2024]),
}

BM25_CASES = {
    "bm25_chunks": dict(company_name=COMPANY, query="gross margin revenue 2024", top_n=5),
    "bm25_pages": dict(company_name=COMPANY, query="cash flow dividend the company", top_n=7, return_parent_pages=True),
    "bm25_years": dict(company_name=COMPANY, query="mexico poland factory capacity", top_n=6, selected_years=[2023, 2025]),
    "bm25_no_match_terms": dict(company_name=OTHER, query="zzzz qqqq", top_n=4),
    "bm25_other": dict(company_name=OTHER, query="tax profit equity debt", top_n=3, selected_years=[2021]),
}

ROUTE_CASES = [
    dict(company_name=COMPANY, question="", fallback_strategy="all", selected_years=None),
    dict(company_name=COMPANY, question="", fallback_strategy="latest", selected_years=None),
    dict(company_name=COMPANY, question="", fallback_strategy="latest", selected_years=[2020, 2022]),
    dict(company_name=COMPANY, question="", fallback_strategy="all", selected_years=[]),
    dict(company_name=COMPANY, question="", fallback_strategy="all", selected_years=[1990]),
    dict(company_name=OTHER, question="", fallback_strategy="latest", selected_years=None),
    dict(company_name="nobody", question="", fallback_strategy="all", selected_years=[2024]),
]

YEAR_QUESTIONS = [
    ("2024年xxx", True), ("2024年相比2023年", True), ("2024年xxx", False), ("no year here", True),
    ("2025年9月30日的现金 2021年 2021年", False), ("2019年和2025年", True),
]


def normalise(obj):
    """JSON-comparable form: tuples -> lists, numpy scalars -> python, timings zeroed."""
    import numpy as np

    if isinstance(obj, dict):
        return {str(k): (0.0 if str(k) in ("hyde_expansion", "multi_query_expansion", "embedding_generation",
                                            "vector_search", "llm_reranking") and isinstance(v, float)
                         else normalise(v)) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [normalise(v) for v in obj]
    if isinstance(obj, (np.floating,)):
        return float(obj)
    if isinstance(obj, (np.integer,)):
        return int(obj)
    return obj
