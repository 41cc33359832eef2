"""ORACLE shim: `compressai` package facade over oracle/compressai_port.py (parity unpinned)."""
