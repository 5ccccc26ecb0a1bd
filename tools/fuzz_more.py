import sys, os, traceback
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo")); sys.path.insert(0, os.path.join(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"), "tests"))
import test_fuzz_gpu as T
bad = []
for seed in range(24, 224):
    try:
        T.test_random_configuration_matches_reference(seed)
    except BaseException as e:
        if type(e).__name__ == "Skipped": continue
        bad.append((seed, repr(e)[:200])); print("FAIL", seed, repr(e)[:300], flush=True)
for seed in range(6, 46):
    try:
        T.test_random_fused_pass_matches_two_reference_passes(seed)
    except BaseException as e:
        if type(e).__name__ == "Skipped": continue
        bad.append(("fused", seed, repr(e)[:200])); print("FAIL fused", seed, repr(e)[:300], flush=True)
print("done; failures:", len(bad), bad[:5])
