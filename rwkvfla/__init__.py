"""Import-compatible stand-in for the `rwkvfla` package (rwkv-fla==0.7.202503140658, third party, not
vendored by the reference) covering exactly the symbols yynil/RWKVTTS imports (SURVEY.md section 8b,
seam 2).  The time-mix recurrence runs in this repository's CUDA library (rwkvtts_b200); there is no
Triton and no CPU path: a forward on CPU tensors raises."""
__version__ = "0.7.202503140658+rwkvtts_b200"
