#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log
tail -8 gpurun_out/m_pytest.log
timeout 900 python tools/time_overloads.py > gpurun_out/m_overloads.jsonl 2> gpurun_out/m_overloads.err; cat gpurun_out/m_overloads.jsonl; tail -2 gpurun_out/m_overloads.err
