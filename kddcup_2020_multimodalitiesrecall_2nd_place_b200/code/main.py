"""code/main.py of the reference: weighted 4-way ensemble, uniqueness filter, top-5 submission (main.py:11-104).
Same file contract (three TSV `qid\\tpid\\tscore`, one CSV with header `query-id,product-id,score`; output CSV with
header `query-id,product1..product5`) and the same default paths, relative to the working directory."""
from ..ensemble import main, merge_and_select, read_scores, write_submission  # noqa: F401

if __name__ == "__main__":
    main()
