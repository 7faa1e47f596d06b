"""Helpers for the tests that run oracle/_ref/nbody_sim_* (the reference's simulation flow with both overlay patches)."""
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINE = re.compile(r"#\s+(\d+)\s+t\s+([0-9.eE+-]+)\s+CC\s+(\d+)\s+SC\s+(\d+)(.*)")


def sim_path(precision="f64"):
    return os.path.join(ROOT, "oracle", "_ref", "nbody_sim_%s" % precision)


def sim_env(extra=None):
    """The statistics lines go through the Qt stand-in's qDebug, which NBREF_QUIET (set for the in-process harness) mutes."""
    env = dict(os.environ, **(extra or {}))
    env.pop("NBREF_QUIET", None)
    return env


def run_sim(precision="f64", timeout=1800, env=None, **opts):
    """Runs the simulation; returns (statistics lines as dicts, the --json summary, stderr text)."""
    cmd = [sim_path(precision), "--json=1"] + ["--%s=%s" % (k, v) for k, v in opts.items()]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=sim_env(env))
    assert res.returncode == 0, res.stderr[-2000:]
    rows = []
    for line in res.stderr.splitlines():
        m = LINE.match(line.strip())
        if not m:
            continue
        row = {"step": int(m.group(1)), "t": float(m.group(2)), "CC": int(m.group(3)), "SC": int(m.group(4))}
        rest = m.group(5).split()
        for k, v in zip(rest[0::2], rest[1::2]):
            if k in ("dP", "dL", "dE", "Vcm"):
                row[k] = float(v)
        rows.append(row)
    summary = json.loads(res.stdout.strip().splitlines()[-1])
    return rows, summary, res.stderr
