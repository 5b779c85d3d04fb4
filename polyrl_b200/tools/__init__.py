"""Command-line utilities around the framework (counterparts of the reference's ``scripts/``: init_random_model.py, print_cfg.py,
diagnose.py, rollout_viewer.py):  ``python -m polyrl_b200.tools.<name> --help``."""
