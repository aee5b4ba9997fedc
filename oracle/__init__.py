"""TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference connector + tools to pin it against the
executing reference.  Nothing under audio-visual-llm_b200/ may import this package: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, and only as the checker."""
