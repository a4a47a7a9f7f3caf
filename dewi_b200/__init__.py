"""Import shim: `import dewi_b200` resolves to the package directory the repository layout
prescribes (its name contains hyphens, so it cannot be imported by name)."""

from pathlib import Path as _Path

_impl = _Path(__file__).resolve().parent.parent / "dewi-design-for-an-entropy-weighted-index-for-text-image-corpora_b200"
__path__ = [str(_impl)]
exec(compile((_impl / "__init__.py").read_text(), str(_impl / "__init__.py"), "exec"))
