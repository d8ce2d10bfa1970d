"""Parity and host-logic tests: `-m "not gpu"` = oracle vs golden vectors of the unmodified reference, host logic, C-ABI symbols;
`-m gpu` = the CUDA path through the C ABI vs the oracle and the golden vectors."""
