/**
 * \file van_factory.h
 * \brief Maps the DMLC_ENABLE_RDMA / PS_VAN_TYPE string to a transport.
 */
#ifndef PS_VAN_VAN_FACTORY_H_
#define PS_VAN_VAN_FACTORY_H_
#include <string>

namespace ps {
class Van;
class Postoffice;
/*! \brief see include/ps/internal/van.h for the accepted names */
Van* CreateVanByType(const std::string& type, Postoffice* postoffice);
}  // namespace ps
#endif  // PS_VAN_VAN_FACTORY_H_
