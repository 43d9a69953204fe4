"""CPU: the reference arm of bench.py (`--impl reference`, the oracle port timed on the host cores) and the bounded CPU
baseline of the bake run end to end on a shrunken workload and print the contract's JSON line."""
import argparse
import json


def test_reference_arm_prints_contract_line(monkeypatch, capsys):
    import bench
    monkeypatch.setattr(bench, "S_TOT", 512)          # the real arm samples S = 9728; same code path, seconds instead of a minute
    args = argparse.Namespace(gpus=1, steps=1, warmup=0, impl="reference", no_cpu_baseline=False, no_bake=False)
    bench.run_reference(args, 0, 1)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["value"] > 0 and line["higher_is_better"] is True and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"] == bench.bench_config(1)      # the identical workload dict the B200 arm prints
    ex = line["extrapolated"]                           # a whole block is measured, the step figure says it is 57 x that
    assert ex["factor"] == 57 and abs(line["ms_per_step"] - 57 * ex["measured_ms"]) < 1e-6 * line["ms_per_step"]
    bench.run_reference(args, 1, 2)                    # other ranks: no work, no output
    assert capsys.readouterr().out == ""


def test_bake_cpu_baseline_runs():
    import bench
    out = bench.bench_uv_bake_cpu()
    assert out["unit"] == "Mpix/s" and out["value"] > 0 and out["kind"] == "port"
