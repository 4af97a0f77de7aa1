#!/usr/bin/env bash
# SURVEY.md §8(c) "extra leg": is R + rstan + bayesynergy available on the GPU box?  If it were, log_prob /
# grad_log_prob at the golden points would pin the oracle.  Writes gpurun_out/r02_rstan_probe.txt.
out=gpurun_out/r02_rstan_probe.txt
mkdir -p gpurun_out
{
  echo "# probe run on $(hostname) at $(date -u +%FT%TZ)"
  echo "## command -v Rscript R"
  command -v Rscript R || echo "(none on PATH)"
  echo "## ls /usr/lib/R /usr/local/lib/R /opt/R"
  ls -d /usr/lib/R /usr/local/lib/R /opt/R /usr/share/R 2>&1
  echo "## find / -name 'libR.so' -o -name 'rstan' -o -name 'StanHeaders' (maxdepth 6)"
  find / -maxdepth 6 \( -name 'libR.so' -o -name 'rstan' -o -name 'StanHeaders' -o -name 'stanc' -o -name 'cmdstan*' \) -not -path '/proc/*' 2>/dev/null | head -20
  echo "(end of find)"
  echo "## Rscript -e 'library(rstan); library(bayesynergy)'"
  if command -v Rscript >/dev/null 2>&1; then Rscript -e 'library(rstan); library(bayesynergy); cat("ok\n")' 2>&1 | tail -20; else echo "Rscript: command not found"; fi
  echo "## python -c 'import cmdstanpy / pystan / rpy2'"
  python - <<'PY'
for m in ("cmdstanpy", "pystan", "stan", "rpy2", "bridgestan"):
    try:
        __import__(m); print(m, "importable")
    except Exception as e:
        print(m, "absent:", type(e).__name__)
PY
  echo "## host"
  nproc; grep -m1 'model name' /proc/cpuinfo
} > "$out" 2>&1
cat "$out"
