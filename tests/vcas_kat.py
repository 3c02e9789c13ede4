"""VerticalCAS known-answer test, restated from /root/reference/models/VerticalCAS/VerticalCAS.jl.

The only result pins in the reference that exercise forward(., net, DeepZ()) without the
Taylor-model plant are the four verdicts asserted at VerticalCAS.jl:311-325:
hdot0 = -19.5, -22.5 -> "verified";  -25.5, -28.5 -> "falsified".

``forward_high(adv_index, c, G) -> high(out)`` is injected so the same loop checks the NumPy
oracle, the C oracle and the CUDA path.
"""
import numpy as np

ADV = ["COC", "DNC", "DND", "DES1500", "CL1500", "SDES1500", "SCL1500", "SDES2500", "SCL2500"]
g = 32.2
ACC_CENTRAL = {"COC": 0.0, "DNC": -7 * g / 24, "DND": 7 * g / 24, "DES1500": -7 * g / 24,
               "CL1500": 7 * g / 24, "SDES1500": -g / 3, "SCL1500": g / 3, "SDES2500": -g / 3,
               "SCL2500": g / 3}
EXPECTED = {-19.5: "verified", -22.5: "verified", -25.5: "falsified", -28.5: "falsified"}


def _complies(adv, lo, hi):
    """hdot (interval [lo, hi], ft/min) subset of advisory2set[adv] (VerticalCAS.jl:71-80)."""
    if adv == "COC":
        return False  # EmptySet
    if adv == "DNC":
        return hi <= 0.0
    if adv == "DND":
        return lo >= 0.0
    if adv in ("DES1500", "SDES1500"):
        return hi <= -1500.0
    if adv in ("CL1500", "SCL1500"):
        return lo >= 1500.0
    if adv == "SDES2500":
        return hi <= -2500.0
    if adv == "SCL2500":
        return lo >= 2500.0
    raise KeyError(adv)


def _extrema(c, G, i):
    r = np.sum(np.abs(G[i, :]))
    return c[i] - r, c[i] + r


def normalize(c, G):
    """normalize(X::LazySet) (VerticalCAS.jl:109-113): translate then diagonal linear map."""
    add = -np.array([0.0, 0.0, 20.0])
    mul = 1.0 / np.array([16000.0, 5000.0, 40.0])
    return np.diag(mul) @ (c + add), np.diag(mul) @ G


def run_instance(hdot0, forward_high, kmax=10):
    """simulate_VerticalCAS + predicate for one initial climb rate; returns (verdict, advisories)."""
    # get_initial_states: convert(Zonotope, Interval(-133,-129) x Singleton([hdot0]))
    c = np.array([-131.0, hdot0])
    G = np.array([[2.0], [0.0]])
    tau, adv = 25.0, "COC"
    A = np.array([[1.0, -1.0], [0.0, 1.0]])
    states = [(c, G)]
    advs = []
    for _ in range(kmax):
        # next_adv: Y = normalize(S x Singleton([tau]))
        cy = np.concatenate([c, [tau]])
        Gy = np.vstack([G, np.zeros((1, G.shape[1]))])
        cy, Gy = normalize(cy, Gy)
        high = forward_high(ADV.index(adv), cy, Gy)
        adv_new = ADV[int(np.argmax(high))]
        # next_acc
        lo2, hi2 = _extrema(c, G, 1)
        comply = _complies(adv_new, 60 * lo2, 60 * hi2)
        hddot = 0.0 if (comply and adv == adv_new) else ACC_CENTRAL[adv_new]
        b = np.array([-hddot * 1.0 ** 2 / 2, hddot * 1.0])
        c, G = A @ c + b, A @ G
        tau -= 1.0
        adv = adv_new
        advs.append(adv_new)
        states.append((c, G))
    # predicate: every projected state is disjoint from {|h| <= 100}
    safe = True
    for (cc, GG) in states:
        lo, hi = _extrema(cc, GG, 0)
        if not (lo > 100.0 or hi < -100.0):
            safe = False
    return ("verified" if safe else "falsified"), advs
