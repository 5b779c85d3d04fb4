#!/bin/bash
# Build (if needed) and run the C++ rollout manager.
python -m polyrl_b200.manager.build >/dev/null
exec polyrl_b200/manager/bin/rollout-manager --config-file ${CONFIG:-examples/configs/manager.toml} --bind-addr ${BIND:-0.0.0.0:5000} "$@"
