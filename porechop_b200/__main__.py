"""`python -m porechop_b200 <porechop arguments>`: run the Porechop CLI found on sys.path (an unmodified checkout whose
`porechop/cpp_functions.so` is this engine's library) with its three alignment phases batched
by `porechop_b200.patch`.  All arguments are Porechop's own (porechop.py:82-214); nothing is added or interpreted here."""
import sys


def main():
    try:
        import porechop
        from porechop import porechop as cli
    except ImportError:
        sys.exit('porechop_b200: no `porechop` package on sys.path (point PYTHONPATH at a Porechop checkout)')
    from . import patch
    memo = patch.install(porechop)
    try:
        cli.main()
    finally:
        patch.uninstall(memo)


if __name__ == '__main__':
    main()
