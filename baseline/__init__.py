"""Reference arm of the benchmark: the UNMODIFIED vendored SuperPoint / LightGlue model files of the reference, staged
(copied, never edited) into the git-ignored ``baseline/_ref/`` and driven the way the reference's plugins drive them."""
