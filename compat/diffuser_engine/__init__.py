"""`diffuser_engine` package name used by the reference's SD pipeline (pipelines/pipeline_stable_diffusion.py:41);
the package itself is not part of the reference checkout - this adapter supplies the one module it imports."""
