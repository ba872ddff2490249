"""Empty stand-in: the reference imports pygame (environment.py:12) but only uses it
for render_mode="human" (environment.py:792-828), which the oracle never selects."""
