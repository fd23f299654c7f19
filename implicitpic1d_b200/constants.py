"""Physical constants, as OpenMP/SharedModules/src/constants.f90:7-15."""
import math

c = 299792458.0
eps_0 = 8.8541878128e-12
m_p = 1.67262192369e-27
m_amu = 1.66053906660e-27
m_e = 9.1093837015e-31
k_B = 1.380649e-23
e = 1.602176634e-19
mu_0 = 1.25663706212e-6
pi = 4.0 * math.atan(1.0)


def getDebyeLength(T_e, n_e):
    """mod_BasicFunctions.f90:16-20"""
    return math.sqrt(eps_0 * T_e / n_e / e)


def getPlasmaFreq(n_e):
    """mod_BasicFunctions.f90:22-26"""
    return math.sqrt(n_e * (e ** 2) / m_e / eps_0)


def del_t_from_deck(n_ave, fractionFreq, del_t_deck):
    """constants.f90:106-107: del_t = MIN(fractionFreq / w_p, del_t)"""
    return min(fractionFreq * 1.0 / getPlasmaFreq(n_ave), del_t_deck)
