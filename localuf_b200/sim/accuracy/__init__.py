from localuf_b200.sim.accuracy.monte_carlo import (  # noqa: F401
    monte_carlo, monte_carlo_results_to_sample_counts, monte_carlo_special, subset_sample)
