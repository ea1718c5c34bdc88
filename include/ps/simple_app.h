/**
 * \file simple_app.h
 * \brief SimpleApp: (int head, string body) request / response between nodes.
 *
 * API parity: reference include/ps/simple_app.h:14-184. Unlike the reference
 * (whose Request CHECK-fails for any receiver other than kServerGroup, SURVEY
 * appendix C) a request may address any node id or group id.
 */
#ifndef PS_SIMPLE_APP_H_
#define PS_SIMPLE_APP_H_
#include <functional>
#include <string>
#include "ps/internal/message.h"
#include "ps/internal/postoffice.h"

namespace ps {

/*! \brief what a SimpleApp handler sees */
struct SimpleData {
  int head;
  std::string body;
  int sender;
  int timestamp;
  int customer_id;
};

class SimpleApp {
 public:
  /*!
   * \param app_id must match between the communicating apps
   * \param customer_id locally unique id of this endpoint
   */
  explicit SimpleApp(int app_id, int customer_id, Postoffice* postoffice = nullptr);
  virtual ~SimpleApp() {
    delete obj_;
    obj_ = nullptr;
  }

  /*! \brief send a request to node / group `recv_id`; returns its timestamp */
  virtual inline int Request(int req_head, const std::string& req_body, int recv_id);
  /*! \brief block until every addressee of `timestamp` has responded */
  virtual inline void Wait(int timestamp) { obj_->WaitRequest(timestamp); }
  /*! \brief answer `recv_req` */
  virtual inline void Response(const SimpleData& recv_req, const std::string& res_body = "");

  using Handle = std::function<void(const SimpleData& recved, SimpleApp* app)>;
  virtual inline void set_request_handle(const Handle& h) {
    CHECK(h) << "invalid request handle";
    request_handle_ = h;
  }
  virtual inline void set_response_handle(const Handle& h) {
    CHECK(h) << "invalid response handle";
    response_handle_ = h;
  }
  virtual inline Customer* get_customer() { return obj_; }

 protected:
  /*! \brief for subclasses that build the Customer themselves */
  inline SimpleApp() : obj_(nullptr), postoffice_(nullptr) {
    request_handle_ = [this](const SimpleData& recved, SimpleApp* app) { app->Response(recved); };
    response_handle_ = [](const SimpleData&, SimpleApp*) {};
  }
  /*! \brief dispatch a received simple_app message to the handlers */
  virtual inline void Process(const Message& msg);

  Customer* obj_;
  Postoffice* postoffice_;

 private:
  Handle request_handle_;
  Handle response_handle_;
};

inline SimpleApp::SimpleApp(int app_id, int customer_id, Postoffice* postoffice) : SimpleApp() {
  using namespace std::placeholders;
  postoffice_ = postoffice ? postoffice : Postoffice::Get();
  obj_ = new Customer(app_id, customer_id, std::bind(&SimpleApp::Process, this, _1), postoffice_);
}

inline int SimpleApp::Request(int req_head, const std::string& req_body, int recv_id) {
  Message msg;
  msg.meta.head = req_head;
  if (!req_body.empty()) msg.meta.body = req_body;
  // group ids open a tracker entry sized to the group; single nodes expect one reply
  const auto& targets = postoffice_->GetNodeIDs(recv_id);
  const int ts = obj_->NewRequest(recv_id, static_cast<int>(targets.size()));
  msg.meta.timestamp = ts;
  msg.meta.request = true;
  msg.meta.simple_app = true;
  msg.meta.app_id = obj_->app_id();
  msg.meta.customer_id = obj_->customer_id();
  for (int r : targets) {
    msg.meta.recver = r;
    postoffice_->van()->Send(msg);
  }
  return ts;
}

inline void SimpleApp::Response(const SimpleData& req, const std::string& res_body) {
  Message msg;
  msg.meta.head = req.head;
  if (!res_body.empty()) msg.meta.body = res_body;
  msg.meta.timestamp = req.timestamp;
  msg.meta.request = false;
  msg.meta.simple_app = true;
  msg.meta.app_id = obj_->app_id();
  msg.meta.customer_id = req.customer_id;
  msg.meta.recver = req.sender;
  postoffice_->van()->Send(msg);
}

inline void SimpleApp::Process(const Message& msg) {
  SimpleData recv;
  recv.sender = msg.meta.sender;
  recv.head = msg.meta.head;
  recv.body = msg.meta.body;
  recv.timestamp = msg.meta.timestamp;
  recv.customer_id = msg.meta.customer_id;
  if (msg.meta.request) {
    CHECK(request_handle_);
    request_handle_(recv, this);
  } else {
    CHECK(response_handle_);
    response_handle_(recv, this);
  }
}

}  // namespace ps
#endif  // PS_SIMPLE_APP_H_
