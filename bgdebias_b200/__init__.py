"""Importable alias of the package directory ``background-debiased-video-cil_b200/``.

The directory name the layout contract asks for contains hyphens, which Python's import
statement cannot spell; this shim makes ``import bgdebias_b200`` load that directory as the
package (its ``__init__.py`` runs below, its sub-modules resolve through ``__path__``).
"""
import pathlib as _pathlib

_real = _pathlib.Path(__file__).resolve().parent.parent / "background-debiased-video-cil_b200"
__path__ = [str(_real)]
exec(compile((_real / "__init__.py").read_text(), str(_real / "__init__.py"), "exec"))
