from localuf_b200.sim.accuracy.monte_carlo import monte_carlo, monte_carlo_results_to_sample_counts  # noqa: F401
