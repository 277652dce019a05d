"""GPU: the two sampler CLIs with the reference's flags (-M -R -f -F -o -t -s --parallel) and the
trajectory analyser, end to end through files."""
import json
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(cmd):
    r = subprocess.run([sys.executable] + cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_ggm_cli_benchpress_and_json(tmp_path):
    from paralleldg_b200 import datagen
    X, _ = datagen.ggm_ar_data(10, 60, seed=2)
    f = tmp_path / "data.csv"
    pd.DataFrame(X, columns=["v%d" % i for i in range(10)]).to_csv(f, index=False)
    out = run([os.path.join("bin", "parallelDG_ggm_sample"), "-M", "400", "-R", "100", "-f", str(f), "-o", str(tmp_path),
               "-F", "traj.csv", "-t", "benchpress", "-s", "7", "--parallel", "--chains", "3"])
    assert "wrote file" in out
    for c in range(3):
        main = pd.read_csv(tmp_path / ("traj_chain%d.csv" % c))
        assert list(main.columns) == ["added", "index", "removed", "score"]
        assert main["index"].tolist()[:3] == [-2, -1, 0]
        assert main["added"].iloc[0].startswith("[v0-v1;")       # column labels are used
        assert (tmp_path / ("traj_chain%d_subindex.csv" % c)).exists()
    # default format: JSON, single chain, single-move kernel (no --parallel), readable by the analyser
    jd = tmp_path / "json"
    run([os.path.join("bin", "parallelDG_ggm_sample"), "-M", "300", "-R", "50", "-f", str(f), "-o", str(jd), "-s", "3"])
    files = [x for x in os.listdir(jd) if x.endswith(".json")]
    assert len(files) == 1 and files[0].startswith("mh_graph_trajectory_ggm_posterior_n_60_p_10")
    js = json.load(open(jd / files[0]))
    assert js["sampling_method"]["method"] == "Metropolis-Hastings - single move sampler"
    assert js["n_updates"] == 300 and len(js["loglikelihood_trace"]) == 300
    run([os.path.join("bin", "analyze_graph_trajectories"), "-i", str(jd), "-o", str(jd / "an"), "--burnin_end", "50"])
    summ = [x for x in os.listdir(jd / "an") if x.endswith("_summary.json")]
    assert len(summ) == 1
    s = json.load(open(jd / "an" / summ[0]))
    assert s["p"] == 10 and 0.0 <= s["acceptance"] <= 1.0
    heat = np.loadtxt(jd / "an" / summ[0].replace("_summary.json", "_heatmap.csv"), delimiter=",")
    assert heat.shape == (10, 10) and np.allclose(heat, heat.T)


def test_loglinear_cli(tmp_path):
    from paralleldg_b200 import datagen
    data, levels, _ = datagen.loglin_ar_data(8, 300, seed=4)
    f = tmp_path / "d.csv"
    cols = pd.MultiIndex.from_arrays([["x%d" % i for i in range(8)], [str(int(l)) for l in levels]])
    pd.DataFrame(data, columns=cols).to_csv(f, index=False)
    run([os.path.join("bin", "parallelDG_loglinear_sample"), "-M", "300", "-R", "100", "-f", str(f), "-o", str(tmp_path),
         "-F", "ll.csv", "-t", "benchpress", "-s", "5", "--parallel"])
    main = pd.read_csv(tmp_path / "ll.csv")
    assert list(main.columns) == ["added", "index", "removed", "score"]
    assert np.isfinite(main["score"].iloc[2])
