"""Print the fully resolved training configuration (defaults + YAML + dotted overrides), validated the way the trainers validate it
(reference: scripts/print_cfg.py).

    python -m polyrl_b200.tools.print_cfg [--config file.yaml] [--json] key=value ...
"""
from __future__ import annotations

import argparse
import json
import sys


def main(argv: list[str] | None = None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config", default=None)
    ap.add_argument("--json", action="store_true")
    ap.add_argument("--no-validate", action="store_true")
    ap.add_argument("overrides", nargs="*")
    a = ap.parse_args(argv)
    from ..config.schema import load_config, validate_config

    cfg = load_config(a.config, overrides=a.overrides)
    if not a.no_validate:
        validate_config(cfg)
    d = cfg.to_dict()
    if a.json:
        print(json.dumps(d, indent=1, default=str))
    else:
        import yaml

        print(yaml.safe_dump(json.loads(json.dumps(d, default=str)), sort_keys=False))
    return 0


if __name__ == "__main__":
    sys.exit(main())
