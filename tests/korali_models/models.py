"""Objective / constraint functions written the way Korali users write them (a Sample-like object with
["Parameters"] in and ["F(x)"] out). Ported from the reference's own test models:
  tests/statistical/optimizers/correctness/model/model.py, detailed/ccmaes/helpers/helpers.py,
  detailed/cmaes/model/model.py, examples/optimization/stochastic/_model/model.py."""
import math


def evalmodel(s):            # correctness/model/model.py:6-9  (maximum 0.23246 at x = -0.45)
    x = s["Parameters"][0]
    s["F(x)"] = -(x * x + math.sin(x))


def constraint1(k):          # correctness/model/model.py:11-12 (can never be satisfied)
    k["F(x)"] = 100.0


def g09(s):                  # examples/optimization/constrained/_model/g09.py (7-D)
    v = s["Parameters"]
    s["F(x)"] = -((v[0] - 10.0) ** 2 + 5.0 * (v[1] - 12.0) ** 2 + v[2] ** 4 + 3.0 * (v[3] - 11.0) ** 2 + 10.0 * v[4] ** 6
                  + 7.0 * v[5] ** 2 + v[6] ** 4 - 4.0 * v[5] * v[6] - 10.0 * v[5] - 8.0 * v[6])


def g1(k):
    v = k["Parameters"]
    k["F(x)"] = -127.0 + 2 * v[0] * v[0] + 3.0 * pow(v[1], 4) + v[2] + 4.0 * v[3] * v[3] + 5.0 * v[4]


def g2(k):
    v = k["Parameters"]
    k["F(x)"] = -282.0 + 7.0 * v[0] + 3.0 * v[1] + 10.0 * v[2] * v[2] + v[3] - v[4]


def g3(k):
    v = k["Parameters"]
    k["F(x)"] = -196.0 + 23.0 * v[0] + v[1] * v[1] + 6.0 * v[5] * v[5] - 8.0 * v[6]


def g4(k):
    v = k["Parameters"]
    k["F(x)"] = 4.0 * v[0] * v[0] + v[1] * v[1] - 3.0 * v[0] * v[1] + 2.0 * v[2] * v[2] + 5.0 * v[5] - 11.0 * v[6]


# ---- detailed/ccmaes/helpers/helpers.py --------------------------------------------------------------------
def evaluateModel(s):
    x1, x2 = s["Parameters"][0], s["Parameters"][1]
    s["F(x)"] = -x1**2 - x2**2 - math.sin(x1)**2 - math.sin(x2)**2


def inactive1(k): k["F(x)"] = -1
def inactive2(k): k["F(x)"] = -2
def activeMax1(k): k["F(x)"] = -(k["Parameters"][0] - 1.0)
def activeMax2(k): k["F(x)"] = -(k["Parameters"][0] - 2.0)
def activeMax3(k): k["F(x)"] = -(k["Parameters"][1] - 1.0)
def activeMax4(k): k["F(x)"] = -(k["Parameters"][1] - 2.0)
def inactiveMax1(k): k["F(x)"] = -math.cos(k["Parameters"][0])
def inactiveMax2(k): k["F(x)"] = -math.sin(k["Parameters"][0])
def inactiveMax3(k): k["F(x)"] = -math.cos(k["Parameters"][1])
def inactiveMax4(k): k["F(x)"] = -math.sin(k["Parameters"][1])
def stress1(k): k["F(x)"] = -k["Parameters"][0] + 6.2
def stress2(k): k["F(x)"] = k["Parameters"][0] - k["Parameters"][1]
def stress3(k): k["F(x)"] = k["Parameters"][0] + 2.0 - 2.0 * k["Parameters"][1]
def stress4(k): k["F(x)"] = 2 * k["Parameters"][0] - 3 * k["Parameters"][1]
def stress5(k): k["F(x)"] = -(k["Parameters"][0] - 6.28) * (k["Parameters"][1] - 6.28)
def stress6(k): k["F(x)"] = -math.cos(k["Parameters"][0]) * math.cos(k["Parameters"][1])
def stress7(k): k["F(x)"] = -math.sin(k["Parameters"][0]) * math.sin(k["Parameters"][1])
def stress8(k): k["F(x)"] = k["Parameters"][0] - k["Parameters"][1]**2


# ---- termination/helpers: 1-D parabola used by cmaes_termination.py --------------------------------------------
def parabola(s):
    x = s["Parameters"][0]
    s["F(x)"] = -x * x


# ---- detailed/cmaes/model/model.py:12-33 ------------------------------------------------------------------------
def make_minmodel(offset):
    def f(s):
        x = s["Parameters"][0]
        s["F(x)"] = -((x - 2.0) * (x - 2.0) + offset)
    return f


def negative_rosenbrock(p):  # examples/optimization/stochastic/_model/model.py:23-34
    x = p["Parameters"]
    res = 0.0
    for i in range(len(x) - 1):
        res += 100 * (x[i + 1] - x[i]**2)**2 + (1 - x[i])**2
    p["F(x)"] = -res


def negative_sphere(p):      # examples/optimization/stochastic/_model/model.py:10-20 (sets "Gradient")
    x = p["Parameters"]
    dim = len(x)
    res = 0.
    grad = [0.] * dim
    for i in range(dim):
        res += x[i]**2
        grad[i] = -x[i]
    p["F(x)"] = -0.5 * res
    p["Gradient"] = grad


def discrete_model(s):       # examples/optimization/discrete/_model/model.py:5-16
    npar = 10
    res = 0.0
    v = s["Parameters"]
    for i in range(npar):
        if (i == 0 or i == 1 or i == 3 or i == 6):
            res += pow(10, 6.0 * i / npar) * round(v[i]) * round(v[i])
        else:
            res += pow(10, 6.0 * i / npar) * v[i] * v[i]
    s["F(x)"] = -res
